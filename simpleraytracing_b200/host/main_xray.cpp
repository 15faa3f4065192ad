// main_xray.cpp -- command-line driver with the flags of the reference's X-ray drivers
// (src/main-pthreads-lbuffer.cxx:427-515, src/main.cxx:269-392):
//   -s/--size W H   image size (default 2048 2048)       -f/--filename NAME  output text file (./out/NAME)
//   -t/--threads N  accepted for script compatibility; the frame is rendered by the GPU
//   -i/--input PLY  mesh file (default ./dragon.ply, as src/main.cxx:194)
//   -g/--gpu ID     CUDA device (extension)               -h/--help
// Flow = main() of main-pthreads-lbuffer.cxx: load -> bbox -> camera -> L-buffer render -> repair -> text file.
#include <chrono>
#include <cstdlib>
#include <exception>
#include <iostream>
#include <string>
#include <vector>

#include "srt/PlyLoader.h"
#include "srt/RenderLoop.h"

static void usage(const char* prog)
{
    std::cerr << "Usage: " << prog << " <option(s)>\nOptions:\n"
              << "\t-h,--help\t\t\tShow this help message\n"
              << "\t-t,--threads T\t\t\tAccepted and ignored (the GPU renders the frame)\n"
              << "\t-s,--size IMG_WIDTH IMG_HEIGHT\tImage size in pixels (default: 2048 2048)\n"
              << "\t-f,--filename NAME\t\tOutput text file, written to ./out/NAME (default: test.txt)\n"
              << "\t-i,--input FILE\t\t\tPLY mesh (default: ./dragon.ply)\n"
              << "\t-g,--gpu ID\t\t\tCUDA device (default: 0)\n";
}

int main(int argc, char** argv)
{
    try {
        std::string out_name = "./out/test.txt", mesh_file = "./dragon.ply";
        unsigned int width = 2048, height = 2048;
        int gpu = 0;
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            auto need = [&](int n) { if (i + n >= argc) { usage(argv[0]); std::exit(EXIT_FAILURE); } };
            if (a == "-h" || a == "--help") { usage(argv[0]); return EXIT_SUCCESS; }
            else if (a == "-s" || a == "--size") { need(2); width = (unsigned)std::stoi(argv[++i]); height = (unsigned)std::stoi(argv[++i]); }
            else if (a == "-f" || a == "--filename") { need(1); out_name = std::string("./out/") + argv[++i]; }
            else if (a == "-t" || a == "--threads") { need(1); ++i; }
            else if (a == "-i" || a == "--input") { need(1); mesh_file = argv[++i]; }
            else if (a == "-g" || a == "--gpu") { need(1); gpu = std::stoi(argv[++i]); }
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
        srt::XRayRenderer renderer(gpu);
        renderer.setMeshes(meshes);
        renderer.setCamera(info, width, height);
        renderer.render(image);
        auto t2 = clock::now();
        const srt_timing tm = renderer.timing();
        std::cout << "Rendering loop took: " << std::chrono::duration<double>(t2 - t1).count() << " seconds (GPU kernels: "
                  << tm.render_ms + tm.repair_ms << " ms, " << tm.pair_tests << " ray-triangle tests)" << std::endl;
        image.saveTextFile(out_name);
        std::cout << "Loading meshes took: " << std::chrono::duration<double>(t1 - t0).count() << " seconds" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
