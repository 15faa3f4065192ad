// main_xray.cpp -- command-line driver with the flags of the reference's X-ray drivers
// (src/main-pthreads-lbuffer.cxx:427-515, src/main.cxx:269-392):
//   -s/--size W H   image size (default 2048 2048)       -f/--filename NAME  output text file (./out/NAME)
//   -t/--threads N  accepted for script compatibility; the frame is rendered by the GPU
//   -i/--input PLY  mesh file (default ./dragon.ply, as src/main.cxx:194)
//   -b/--background R G B   parsed and range-checked like src/main.cxx:324-365; the X-ray image is a float image without
//                   a background colour (the reference's main() does not use the values either)
//   -g/--gpu ID     CUDA device (extension)               -G/--gpus N  render on the first N GPUs (extension)
//   --csv FILE      append a row "CPU,Parallelisation,Number of threads,Number of nodes,Compiler,Image size,Runtime in
//                   sec" in the format of the reference's batch scripts (submit-serial.sh:15-22)      -h/--help
// Flow = main() of main-pthreads-lbuffer.cxx: load -> bbox -> camera -> L-buffer render -> repair -> text file.
#include <chrono>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "srt/PlyLoader.h"
#include "srt/RenderLoop.h"

#ifndef SRT_NVCC_VERSION
#define SRT_NVCC_VERSION "12.9"
#endif

static void usage(const char* prog)
{
    std::cerr << "Usage: " << prog << " <option(s)>\nOptions:\n"
              << "\t-h,--help\t\t\tShow this help message\n"
              << "\t-t,--threads T\t\t\tAccepted and ignored (the GPU renders the frame)\n"
              << "\t-s,--size IMG_WIDTH IMG_HEIGHT\tImage size in pixels (default: 2048 2048)\n"
              << "\t-f,--filename NAME\t\tOutput text file, written to ./out/NAME (default: test.txt)\n"
              << "\t-i,--input FILE\t\t\tPLY mesh (default: ./dragon.ply)\n"
              << "\t-b,--background R G B\t\tBackground colour, values between 0 and 255 (accepted; the X-ray image has none)\n"
              << "\t-g,--gpu ID\t\t\tCUDA device (default: 0)\n"
              << "\t-G,--gpus N\t\t\tRender on the first N GPUs of this machine (default: 1)\n"
              << "\t--csv FILE\t\t\tAppend a timing row in the format of the reference's submit-*.sh scripts\n";
}

int main(int argc, char** argv)
{
    try {
        std::string out_name = "./out/test.txt", mesh_file = "./dragon.ply";
        unsigned int width = 2048, height = 2048;
        int gpu = 0, n_gpus = 1;
        std::string csv_file;
        const auto t_start = std::chrono::high_resolution_clock::now();
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            auto need = [&](int n) { if (i + n >= argc) { usage(argv[0]); std::exit(EXIT_FAILURE); } };
            if (a == "-h" || a == "--help") { usage(argv[0]); return EXIT_SUCCESS; }
            else if (a == "-s" || a == "--size") { need(2); width = (unsigned)std::stoi(argv[++i]); height = (unsigned)std::stoi(argv[++i]); }
            else if (a == "-f" || a == "--filename") { need(1); out_name = std::string("./out/") + argv[++i]; }
            else if (a == "-t" || a == "--threads") { need(1); ++i; }
            else if (a == "-i" || a == "--input") { need(1); mesh_file = argv[++i]; }
            else if (a == "-g" || a == "--gpu") { need(1); gpu = std::stoi(argv[++i]); }
            else if (a == "-G" || a == "--gpus") { need(1); n_gpus = std::stoi(argv[++i]); }
            else if (a == "--csv") { need(1); csv_file = argv[++i]; }
            else if (a == "-b" || a == "--background") {
                need(3);
                for (int k = 0; k < 3; ++k) {
                    const int v = std::stoi(argv[++i]);
                    if (v < 0 || v > 255) { usage(argv[0]); return EXIT_FAILURE; }
                }
            }
            else { usage(argv[0]); return EXIT_FAILURE; }
        }
        using clock = std::chrono::high_resolution_clock;
        auto t0 = clock::now();
        std::cout << "Loading polygon meshes" << std::endl;
        std::vector<srt::TriangleMesh> meshes;
        srt::loadMeshes(mesh_file, meshes);
        std::cout << "Retreiving scenes bbox" << std::endl;
        srt::Vec3 lower, upper;
        srt::getBBox(meshes, upper, lower);
        srt::Image image;
        srt::RayTracerInfo info = srt::initialiseRayTracing(meshes, upper, lower, height, width, image, 0.0f);
        auto t1 = clock::now();
        if (n_gpus > 1) {
            std::vector<int> devices(n_gpus);
            for (int i = 0; i < n_gpus; ++i) devices[i] = i;
            srt::MultiGpuXRayRenderer renderer(devices);
            renderer.setMeshes(meshes);
            renderer.setCamera(info, width, height);
            renderer.render(image);
            std::cout << "Rendering loop took: " << std::chrono::duration<double>(clock::now() - t1).count() << " seconds ("
                      << n_gpus << " GPUs, device step: " << renderer.stepMilliseconds() << " ms)" << std::endl;
        } else {
            srt::XRayRenderer renderer(gpu);
            renderer.setMeshes(meshes);
            renderer.setCamera(info, width, height);
            renderer.render(image);
            const srt_timing tm = renderer.timing();
            std::cout << "Rendering loop took: " << std::chrono::duration<double>(clock::now() - t1).count() << " seconds (GPU kernels: "
                      << tm.render_ms + tm.repair_ms << " ms, " << tm.pair_tests << " ray-triangle tests)" << std::endl;
        }
        image.saveTextFile(out_name);
        if (!csv_file.empty()) {
            // CPU,Parallelisation,Number of threads,Number of nodes,Compiler,Image size,Runtime in sec  (submit-serial.sh:15-22;
            // runtime = whole process wall clock like /usr/bin/time %e; "CPU" names the device that did the work)
            bool fresh = !std::ifstream(csv_file).good();
            std::ofstream csv(csv_file, std::ios::app);
            if (fresh) csv << "CPU,Parallelisation,Number of threads,Number of nodes,Compiler,Image size,Runtime in sec\n";
            const double secs = std::chrono::duration<double>(clock::now() - t_start).count();
            csv << "NVIDIA B200 (sm_100a),CUDA," << n_gpus << ",1,nvcc " << SRT_NVCC_VERSION << "," << width << "x" << height << "," << secs << "\n";
        }
        std::cout << "Loading meshes took: " << std::chrono::duration<double>(t1 - t0).count() << " seconds" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
