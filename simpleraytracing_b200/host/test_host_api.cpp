// test_host_api.cpp -- checks of the C++ host classes (include/srt/*.h) against the reference's documented
// behaviour.  `test_host_api cpu` needs no GPU; `test_host_api gpu <ply> <W> <H> <out.txt>` renders through
// srt::renderLoop (CUDA; main.cxx's sort-based image, or the L-buffer flow with a trailing "lbuffer") and writes the image
// with Image::saveTextFile; `test_host_api gpu-multi <ply> <W> <H> <out.txt>
// [n_gpus]` does the same on all (or the first n) visible GPUs through srt::MultiGpuXRayRenderer and checks that the image
// equals the single-GPU one bit for bit.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "srt/PlyLoader.h"
#include "srt/RenderLoop.h"
#include "../csrc/host_copy.h"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)
template <class E, class F> static bool throws(F f) { try { f(); } catch (const E&) { return true; } catch (...) { return false; } return false; }

static int cpu_tests()
{
    using namespace srt;
    // Vec3: float vs double overloads (Vec3.inl:205-233)
    Vec3 a(1.0f, 2.0f, 3.0f), b(0.1f, 0.2f, 0.3f);
    CHECK((a + b).getX() == 1.0f + 0.1f && (a - b).getZ() == 3.0f - 0.3f);
    CHECK((b * 0.6).getX() == float(double(0.1f) * 0.6));
    CHECK((b * 0.6f).getX() == 0.1f * 0.6f);
    CHECK((b / 3.0).getY() == float(double(0.2f) / 3.0) && (b / 3.0f).getY() == 0.2f / 3.0f);
    CHECK(a.dotProduct(b) == 1.0f * 0.1f + 2.0f * 0.2f + 3.0f * 0.3f);
    CHECK(a.crossProduct(b).getX() == 2.0f * 0.3f - 3.0f * 0.2f);
    CHECK(a[2] == 3.0f && throws<std::out_of_range>([&] { (void)a[3]; }));
    Vec3 n(3, 4, 0); n.normalise();
    CHECK(n.getX() == 3.0f / 5.0f && n.getY() == 4.0f / 5.0f && Vec3(3, 4, 0).getLength() == 5.0f);
    CHECK(reflect(Vec3(1, -1, 0), Vec3(0, 1, 0)).getY() == 1.0f);
    std::ostringstream os; os << Vec3(1, 2.5f, -3);
    CHECK(os.str() == "1 2.5 -3");
    // Triangle: cached unit normal, area
    Triangle t(Vec3(0, 0, 0), Vec3(2, 0, 0), Vec3(0, 2, 0));
    CHECK(t.getNormal().getZ() == 1.0f && t.getNormal().getX() == 0.0f && t.getArea() == 2.0f);
    t.setVertices(Vec3(0, 0, 0), Vec3(0, 2, 0), Vec3(2, 0, 0));
    CHECK(t.getNormal().getZ() == -1.0f);
    // Ray: constructor renormalises, zero direction stays zero; intersect is two-sided and accepts negative t
    Ray r(Vec3(0.5f, 0.5f, -1), Vec3(0, 0, 10));
    CHECK(r.getDirection().getZ() == 1.0f);
    CHECK(Ray(Vec3(), Vec3()).getDirection().getLength() == 0.0f);
    float tt = 0;
    CHECK(r.intersect(t, tt) && tt == 1.0f);
    CHECK(Ray(Vec3(0.5f, 0.5f, 1), Vec3(0, 0, 1)).intersect(t, tt) && tt == -1.0f);
    CHECK(!Ray(Vec3(3, 3, -1), Vec3(0, 0, 1)).intersect(t, tt));
    CHECK(r.getPointAt(2.0f).getZ() == 1.0f);
    // TriangleMesh: overloads, length errors, bbox, stubbed bbox test
    std::vector<float> v = { 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 2 };
    std::vector<unsigned int> idx = { 0, 1, 2, 0, 2, 3 };
    TriangleMesh m(v, idx);
    CHECK(m.getNumberOfTriangles() == 2 && m.getUpperBBoxCorner().getZ() == 2.0f && m.getLowerBBoxCorner().getX() == 0.0f);
    CHECK(m.getTriangle(1).getP3().getZ() == 2.0f && m.intersectBBox(r));
    CHECK(throws<std::length_error>([&] { TriangleMesh bad(std::vector<float>(10, 0.f)); }));
    CHECK(throws<std::length_error>([&] { TriangleMesh bad(v, std::vector<unsigned int>{ 0, 1 }); }));
    CHECK(throws<std::length_error>([&] { TriangleMesh bad(v, idx, std::vector<float>(3, 0.f)); }));
    TriangleMesh soup(std::vector<float>{ 0, 0, 0, 1, 0, 0, 0, 1, 0 });
    CHECK(soup.getNumberOfTriangles() == 1 && soup.flatVertices().size() == 9 && soup.flatNormals()[2] == 1.0f);
    std::vector<float> tc = { 0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 0, 0 };
    TriangleMesh tex(v, idx, tc);
    CHECK(tex.getTriangle(1).getTextCoord2().getX() == 1.0f && tex.getTriangle(1).getTextCoord3().getX() == 0.0f);
    // Image: bounds checks, text format, LUT
    Image img(3, 2, 80.0f);
    img.setPixel(1, 0, 11.513f); img.setPixel(2, 1, 0.5f);
    float px = 0; img.getPixel(1, 0, px);
    CHECK(px == 11.513f && throws<std::out_of_range>([&] { img.setPixel(3, 0, 1.0f); }) &&
          throws<std::out_of_range>([&] { img.getPixel(0, 2, px); }));
    const std::string path = "/tmp/srt_test_image.txt";
    img.saveTextFile(path);
    { std::ifstream f(path); std::stringstream ss; ss << f.rdbuf(); CHECK(ss.str() == "80\t11.513\t80\n80\t80\t0.5"); }
    Image copy = img; copy.setPixel(0, 0, 1.0f); img.getPixel(0, 0, px);
    CHECK(px == 80.0f);                                                  // deep copy
    const std::vector<unsigned char> lut = img.applyLUT(0.0f, 80.0f);
    CHECK(lut.size() == 18 && lut[0] == 255 && lut[3] == (unsigned char)std::lround(255.0 * 11.513f / 80.0) && lut[15] == 2);
    // ImageRGB (old Image API): RGB pixels, bounds checks, PPM round trip
    ImageRGB rgb(4, 3, 128, 128, 128);
    rgb.setPixel(3, 2, 1, 2, 3);
    unsigned char cr, cg, cb; rgb.getPixel(3, 2, cr, cg, cb);
    CHECK(cr == 1 && cg == 2 && cb == 3 && rgb.getData()[(2 * 4 + 3) * 3 + 2] == 3);
    CHECK(throws<std::out_of_range>([&] { rgb.setPixel(4, 0, 0, 0, 0); }));
    rgb.savePPMFile("/tmp/srt_test_image.ppm");
    ImageRGB back("/tmp/srt_test_image.ppm"); back.getPixel(3, 2, cr, cg, cb);
    CHECK(back.getWidth() == 4 && back.getHeight() == 3 && cb == 3);
    TriangleMesh bgm = createBackground(Vec3(1, 1, 1), Vec3(0, 0, 0), rgb);
    CHECK(bgm.getNumberOfTriangles() == 2 && bgm.getTextureRGB().getWidth() == 4 && bgm.getTriangle(0).getP1().getX() == 1.0f + 1.0f * 0.1f);
    CHECK(bgm.getTriangle(1).getTextCoord2().getX() == 1.0f && bgm.flatTexCoords().size() == 18);
    // host helpers of the C ABI through the C++ shim
    std::vector<TriangleMesh> set = { m };
    Vec3 lo, hi; getBBox(set, hi, lo);
    CHECK(hi.getZ() == 2.0f && lo.getY() == 0.0f);
    Image out;
    RayTracerInfo info = initialiseRayTracing(set, hi, lo, 8, 8, out, 0.0f);
    CHECK(out.getWidth() == 8 && info.up.getZ() == -1.0f && info.right.getY() == 1.0f && info.range.getZ() == 2.0f);
    CHECK(info.detector_position.getX() > info.origin.getX());
    // no GPU -> the renderer must fail loudly
    int ndev = 0;
    {
        srt_ctx* c = nullptr;
        const int rc = srt_create(&c, 0);
        if (rc == SRT_OK) { srt_destroy(c); ndev = 1; }
    }
    if (!ndev) CHECK(throws<std::runtime_error>([&] { renderLoop(out, set, info); }));
    if (!ndev) CHECK(throws<std::runtime_error>([&] { renderLoopLBuffer(out, set, info); }));
    std::printf(failures ? "%d FAILURES\n" : "OK\n", failures);
    return failures ? 1 : 0;
}

// `parity <in.bin> <out.bin>`: records of 15 floats (ray origin, direction, triangle) -> 9 floats each: the Ray's stored
// direction, hit flag, t, the Triangle's cached normal and its area.  tests/test_cpp_host.py compares them bit for bit with
// what the reference's own Ray / Triangle classes compute for the same records.
static int parity_dump(const char* in_path, const char* out_path)
{
    using namespace srt;
    std::FILE* in = std::fopen(in_path, "rb");
    std::FILE* out = std::fopen(out_path, "wb");
    if (!in || !out) { std::printf("ERROR: cannot open files\n"); return 1; }
    float r[15];
    size_t n = 0;
    while (std::fread(r, sizeof(float), 15, in) == 15) {
        const Ray ray(Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]));
        const Triangle tri(Vec3(r[6], r[7], r[8]), Vec3(r[9], r[10], r[11]), Vec3(r[12], r[13], r[14]));
        float t = 0.0f;
        const bool hit = ray.intersect(tri, t);
        const float o[9] = { ray.getDirection().getX(), ray.getDirection().getY(), ray.getDirection().getZ(), hit ? 1.0f : 0.0f,
                             hit ? t : 0.0f, tri.getNormal().getX(), tri.getNormal().getY(), tri.getNormal().getZ(), tri.getArea() };
        std::fwrite(o, sizeof(float), 9, out);
        ++n;
    }
    std::fclose(in); std::fclose(out);
    std::printf("OK %zu records\n", n);
    return 0;
}

// `pool`: one HostCopyPool reused for many frames like a context reuses its pool (activate -> chunks trickle in -> drain,
// several rounds per frame), for every thread count and polling time; every byte must arrive and none beyond.
static int pool_tests()
{
    uint32_t state = 12345u;
    auto rnd = [&state] { state = state * 1664525u + 1013904223u; return state >> 8; };
    const size_t N = 3u << 20;
    std::vector<unsigned char> src(N), dst(N);
    for (auto& v : src) v = (unsigned char)rnd();
    for (unsigned spin : { 0u, 20u, 2000u })
        for (int threads : { 0, 1, 3, 7 }) {
            srt::HostCopyPool pool(threads, spin);
            for (int it = 0; it < 120; ++it) {
                const size_t n = rnd() % N, chunk = 1 + rnd() % (1u << 18);
                std::fill(dst.begin(), dst.end(), 0);
                const int rounds = 1 + (int)(rnd() % 3);
                size_t done = 0;
                for (int r = 0; r < rounds; ++r) {
                    const size_t part = (r == rounds - 1) ? n - done : (n - done) / 2;
                    pool.activate();
                    if (it % 7 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rnd() % 300));
                    for (size_t off = 0; off < part; off += chunk)
                        pool.submit(dst.data() + done + off, src.data() + done + off, std::min(chunk, part - off));
                    pool.drain();
                    done += part;
                }
                CHECK(std::memcmp(dst.data(), src.data(), n) == 0);
                bool clean = true;
                for (size_t i = n; i < std::min(N, n + 64); ++i) clean = clean && dst[i] == 0;
                CHECK(clean);
            }
        }
    std::printf(failures ? "%d FAILURES\n" : "OK\n", failures);
    return failures ? 1 : 0;
}

int main(int argc, char** argv)
{
    if (argc >= 2 && std::strcmp(argv[1], "cpu") == 0) return cpu_tests();
    if (argc >= 2 && std::strcmp(argv[1], "pool") == 0) return pool_tests();
    if (argc >= 4 && std::strcmp(argv[1], "parity") == 0) return parity_dump(argv[2], argv[3]);
    if (argc >= 6 && std::strcmp(argv[1], "gpu") == 0) {
        try {
            std::vector<srt::TriangleMesh> meshes;
            srt::loadMeshes(argv[2], meshes);
            const unsigned W = (unsigned)std::atoi(argv[3]), H = (unsigned)std::atoi(argv[4]);
            srt::Vec3 lo, hi; srt::getBBox(meshes, hi, lo);
            srt::Image image;
            srt::RayTracerInfo info = srt::initialiseRayTracing(meshes, hi, lo, H, W, image, 0.0f);
            // argv[6] == "lbuffer": the main-pthreads-lbuffer flow; default: main.cxx's sort-based renderLoop
            if (argc >= 7 && std::strcmp(argv[6], "lbuffer") == 0) srt::renderLoopLBuffer(image, meshes, info);
            else srt::renderLoop(image, meshes, info);
            image.saveTextFile(argv[5]);
            std::printf("OK %zu triangles\n", meshes[0].getNumberOfTriangles());
            return 0;
        } catch (const std::exception& e) { std::printf("ERROR: %s\n", e.what()); return 1; }
    }
    if (argc >= 6 && std::strcmp(argv[1], "gpu-multi") == 0) {
        try {
            std::vector<srt::TriangleMesh> meshes;
            srt::loadMeshes(argv[2], meshes);
            const unsigned W = (unsigned)std::atoi(argv[3]), H = (unsigned)std::atoi(argv[4]);
            srt::Vec3 lo, hi; srt::getBBox(meshes, hi, lo);
            srt::Image single, multi;
            srt::RayTracerInfo info = srt::initialiseRayTracing(meshes, hi, lo, H, W, single, 0.0f);
            srt::renderLoopLBuffer(single, meshes, info);                        // GPU 0 alone
            int n = 0;
            while (n < 8) { srt_ctx* c = nullptr; if (srt_create(&c, n) != SRT_OK) break; srt_destroy(c); ++n; }
            if (argc >= 7) n = std::min(n, std::atoi(argv[6]));
            if (n < 1) { std::printf("ERROR: no GPU\n"); return 1; }
            std::vector<int> devices(n);
            for (int i = 0; i < n; ++i) devices[i] = i;
            srt::MultiGpuXRayRenderer r(devices);
            r.setMeshes(meshes);
            r.setCamera(info, W, H);
            for (int rep = 0; rep < 3; ++rep) r.render(multi);                   // both frame sets get used
            const float ms = r.stepMilliseconds();
            size_t diff = 0;
            for (size_t i = 0; i < (size_t)W * H; ++i)
                diff += std::memcmp(single.getData() + i, multi.getData() + i, sizeof(float)) != 0;
            multi.saveTextFile(argv[5]);
            std::printf("%s %d GPUs, %zu differing pixels, step %.3f ms\n", diff ? "MISMATCH" : "OK", n, diff, ms);
            return diff ? 1 : 0;
        } catch (const std::exception& e) { std::printf("ERROR: %s\n", e.what()); return 1; }
    }
    if (argc >= 3 && std::strcmp(argv[1], "ply") == 0) {
        try {
            std::vector<srt::TriangleMesh> meshes;
            srt::loadMeshes(argv[2], meshes);
            const std::vector<float> v = meshes[0].flatVertices();
            double sum = 0, wsum = 0;
            for (size_t i = 0; i < v.size(); ++i) { sum += v[i]; wsum += v[i] * (double)((i % 97) + 1); }
            std::printf("%zu %.17g %.17g\n", meshes[0].getNumberOfTriangles(), sum, wsum);
            return 0;
        } catch (const std::exception& e) { std::printf("ERROR: %s\n", e.what()); return 1; }
    }
    if (argc >= 7 && std::strcmp(argv[1], "shaded") == 0) {
        // main-omp.cxx flow: load, material override (:197), bbox, camera, light (:235), background quad, renderLoop
        try {
            using namespace srt;
            std::vector<TriangleMesh> meshes;
            loadMeshes(argv[2], meshes);
            meshes[0].setMaterial(Material(0.2 * Vec3(1, 0, 0), Vec3(0, 1, 0), Vec3(0, 0, 1), 1));
            const unsigned W = (unsigned)std::atoi(argv[3]), H = (unsigned)std::atoi(argv[4]);
            Vec3 lo, hi; getBBox(meshes, hi, lo);
            Image unused;
            RayTracerInfo info = initialiseRayTracing(meshes, hi, lo, H, W, unused, 0.0f);
            const Vec3 light_position = info.origin + info.up * 100.0;
            Light light(Vec3(1, 1, 1), (lo + info.range / 2.0) - light_position, light_position);
            meshes.push_back(createBackground(hi, lo, ImageRGB(argv[5])));
            ImageRGB out(W, H, 128, 128, 128);
            renderLoop(out, meshes, info.detector_position, info.origin, info.up, info.right, light);
            out.savePPMFile(argv[6]);
            std::printf("OK\n");
            return 0;
        } catch (const std::exception& e) { std::printf("ERROR: %s\n", e.what()); return 1; }
    }
    std::printf("usage: %s cpu | pool | parity <in.bin> <out.bin> | ply <file> | gpu <ply> <W> <H> <out.txt> | shaded <ply> <W> <H> <tex.ppm> <out.ppm>\n", argv[0]);
    return 2;
}
