// srt/RenderLoop.h -- the reference's driver-level render functions on top of the C ABI (include/srt.h).
//
//   reference                                                     here
//   getBBox(meshes, upper, lower)       lbuffer.cxx:585-608       srt::getBBox            -> srt_scene_bbox
//   initialiseRayTracing(...)           lbuffer.cxx:611-667       srt::initialiseRayTracing -> srt_camera_from_bbox
//   renderLoop(Image&, meshes, info)    main.cxx:626-743          srt::renderLoop         -> srt_render_xray_sorted
//   main() of the L-buffer driver       lbuffer.cxx:258-423       srt::renderLoopLBuffer  -> srt_render_xray (L-buffer + repair)
//   renderLoopCallBack(PThreadData*)    lbuffer.cxx:733-813       srt::XRayRenderer::renderRows (row band)
//   repair + setPixel loop in main()    lbuffer.cxx:329-404       srt::XRayRenderer::repair
//   MPI partition + gather              main-mpi.cxx:258-293,854-881   srt::MultiGpuXRayRenderer (one process, N GPUs)
//
// All rendering happens in the CUDA kernels of libsrt_b200.so. Failures (no GPU, CUDA errors, bad
// arguments) throw std::runtime_error carrying srt_last_error(); nothing here computes a frame on the CPU.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../srt.h"
#include "Image.h"
#include "ImageRGB.h"
#include "Light.h"
#include "TriangleMesh.h"
#include "Vec3.h"

namespace srt {

// main-pthreads-lbuffer.cxx:125-135
struct RayTracerInfo {
    Vec3 detector_position;
    Vec3 origin;
    Vec3 up;
    Vec3 right;
    Light light;
    Vec3 upper_bbox_corner;
    Vec3 lower_bbox_corner;
    Vec3 range;
};

namespace detail {
inline void check(int rc, const srt_ctx* ctx, const char* what)
{
    if (rc != SRT_OK) throw std::runtime_error(std::string(what) + ": " + srt_last_error(ctx));
}
struct FlatMeshes {
    std::vector<std::vector<float>> vertices, normals, texcoords;
    std::vector<srt_mesh_desc> desc;
};
// mu: 0.1037f for mesh 0 ("soft tissue"), 0.3971f for the others ("bone"), lbuffer.cxx:798-803
inline FlatMeshes flatten(const std::vector<TriangleMesh>& meshes, float mu0 = 0.1037f, float mu_other = 0.3971f)
{
    FlatMeshes f;
    f.vertices.reserve(meshes.size()); f.normals.reserve(meshes.size()); f.texcoords.reserve(meshes.size());
    for (const TriangleMesh& m : meshes) {
        f.vertices.push_back(m.flatVertices()); f.normals.push_back(m.flatNormals()); f.texcoords.push_back(m.flatTexCoords());
    }
    for (size_t i = 0; i < meshes.size(); ++i) {
        srt_mesh_desc d{};
        d.vertices = f.vertices[i].data();
        d.normals = f.normals[i].data();
        d.n_triangles = (uint32_t)meshes[i].getNumberOfTriangles();
        d.mu = i == 0 ? mu0 : mu_other;
        const Material& mat = meshes[i].getMaterial();
        for (int k = 0; k < 3; ++k) {
            d.material.ambient[k] = mat.getAmbient()[k]; d.material.diffuse[k] = mat.getDiffuse()[k];
            d.material.specular[k] = mat.getSpecular()[k];
        }
        d.material.shininess = mat.getShininess();
        const ImageRGB& tex = meshes[i].getTextureRGB();
        if (tex.getWidth() > 0 && tex.getHeight() > 0) {
            d.texcoords = f.texcoords[i].data();
            d.texture_rgb = tex.getData(); d.tex_w = tex.getWidth(); d.tex_h = tex.getHeight();
        }
        f.desc.push_back(d);
    }
    return f;
}
} // namespace detail

inline void getBBox(const std::vector<TriangleMesh>& meshes, Vec3& upper, Vec3& lower)
{
    const detail::FlatMeshes f = detail::flatten(meshes);
    float lo[3], hi[3];
    detail::check(srt_scene_bbox(f.desc.data(), (int)f.desc.size(), lo, hi), nullptr, "srt_scene_bbox");
    lower = Vec3(lo[0], lo[1], lo[2]); upper = Vec3(hi[0], hi[1], hi[2]);
}

inline RayTracerInfo initialiseRayTracing(std::vector<TriangleMesh>& /*meshes*/, const Vec3& upper, const Vec3& lower,
                                          unsigned int image_height, unsigned int image_width, Image& output_image,
                                          float lut)
{
    srt_camera cam{};
    detail::check(srt_camera_from_bbox(lower.data(), upper.data(), image_width, image_height, &cam), nullptr,
                  "srt_camera_from_bbox");
    output_image = Image(image_width, image_height, lut);                    // lbuffer.cxx:641
    RayTracerInfo info;
    info.detector_position = Vec3(cam.detector[0], cam.detector[1], cam.detector[2]);
    info.origin = Vec3(cam.origin[0], cam.origin[1], cam.origin[2]);
    info.up = Vec3(cam.up[0], cam.up[1], cam.up[2]);
    info.right = Vec3(cam.right[0], cam.right[1], cam.right[2]);
    info.upper_bbox_corner = upper; info.lower_bbox_corner = lower;
    info.range = upper - lower;
    const Vec3 centre = lower + info.range / 2.0;
    info.light = Light(Vec3(1, 1, 1), centre - info.origin, info.origin);   // lbuffer.cxx:643-646
    return info;
}

// RAII owner of one GPU context.
class XRayRenderer {
public:
    explicit XRayRenderer(int device = 0)
    {
        if (srt_create(&ctx_, device) != SRT_OK) throw std::runtime_error(std::string("srt_create: ") + srt_last_error(nullptr));
        srt_default_params(&params_);
    }
    ~XRayRenderer() { srt_destroy(ctx_); }
    XRayRenderer(const XRayRenderer&) = delete;
    XRayRenderer& operator=(const XRayRenderer&) = delete;

    srt_params& params() { return params_; }
    srt_ctx* handle() { return ctx_; }

    void setMeshes(const std::vector<TriangleMesh>& meshes)
    {
        const detail::FlatMeshes f = detail::flatten(meshes);
        detail::check(srt_set_meshes(ctx_, (int)f.desc.size(), f.desc.data()), ctx_, "srt_set_meshes");
        detail::check(srt_synchronize(ctx_), ctx_, "srt_synchronize");     // the flattened arrays die with this scope
    }

    void setCamera(const RayTracerInfo& info, unsigned int width, unsigned int height)
    {
        for (int k = 0; k < 3; ++k) {
            cam_.origin[k] = info.origin[k]; cam_.detector[k] = info.detector_position[k];
            cam_.up[k] = info.up[k]; cam_.right[k] = info.right[k];
        }
        // lbuffer.cxx:739-741
        const float res1 = info.range[2] / width, res2 = info.range[1] / height;
        cam_.pixel_spacing[0] = cam_.pixel_spacing[1] = 2 * std::max(res1, res2);
        cam_.width = width; cam_.height = height;
        detail::check(srt_set_camera(ctx_, &cam_), ctx_, "srt_set_camera");
    }

    // rows [row_begin, row_end) of the pre-repair L-buffer into a full-frame array (the reference's L_buffer)
    void renderRows(unsigned int row_begin, unsigned int row_end, float* L_buffer)
    {
        detail::check(srt_render_lbuffer(ctx_, &params_, row_begin, row_end, L_buffer, nullptr, nullptr, nullptr), ctx_,
                      "srt_render_lbuffer");
    }
    void repair(const float* L_buffer, Image& out)
    {
        out.setSize(cam_.width, cam_.height);
        detail::check(srt_repair(ctx_, &params_, L_buffer, out.getData()), ctx_, "srt_repair");
    }
    // whole frame: L-buffer, repair, image
    void render(Image& out)
    {
        out.setSize(cam_.width, cam_.height);
        detail::check(srt_render_xray(ctx_, &params_, out.getData()), ctx_, "srt_render_xray");
    }
    srt_timing timing() { srt_timing t{}; srt_last_timing(ctx_, &t); return t; }

private:
    srt_ctx* ctx_ = nullptr;
    srt_params params_{};
    srt_camera cam_{};
};

// One process, several GPUs: the role of the MPI driver (src/main-mpi.cxx:258-293 partition, :854-881 gather).  Every
// GPU renders its share of 16-row blocks and stores it into the frame of every GPU over NVLink; render() returns the
// repaired image of the first device, bit-identical to XRayRenderer::render.
class MultiGpuXRayRenderer {
public:
    explicit MultiGpuXRayRenderer(const std::vector<int>& devices)
    {
        if (srt_multi_create(&m_, devices.data(), (int)devices.size()) != SRT_OK)
            throw std::runtime_error(std::string("srt_multi_create: ") + srt_multi_last_error(nullptr));
        srt_default_params(&params_);
    }
    ~MultiGpuXRayRenderer() { srt_multi_destroy(m_); }
    MultiGpuXRayRenderer(const MultiGpuXRayRenderer&) = delete;
    MultiGpuXRayRenderer& operator=(const MultiGpuXRayRenderer&) = delete;

    srt_params& params() { return params_; }
    int deviceCount() const { return srt_multi_device_count(m_); }

    void setMeshes(const std::vector<TriangleMesh>& meshes)
    {
        const detail::FlatMeshes f = detail::flatten(meshes);
        check(srt_multi_set_meshes(m_, (int)f.desc.size(), f.desc.data()), "srt_multi_set_meshes");
        check(srt_multi_synchronize(m_), "srt_multi_synchronize");           // the flattened arrays die with this scope
    }
    void setCamera(const RayTracerInfo& info, unsigned int width, unsigned int height)
    {
        for (int k = 0; k < 3; ++k) {
            cam_.origin[k] = info.origin[k]; cam_.detector[k] = info.detector_position[k];
            cam_.up[k] = info.up[k]; cam_.right[k] = info.right[k];
        }
        const float res1 = info.range[2] / width, res2 = info.range[1] / height;      // lbuffer.cxx:739-741
        cam_.pixel_spacing[0] = cam_.pixel_spacing[1] = 2 * std::max(res1, res2);
        cam_.width = width; cam_.height = height;
        check(srt_multi_set_camera(m_, &cam_), "srt_multi_set_camera");
    }
    void render(Image& out)
    {
        out.setSize(cam_.width, cam_.height);
        check(srt_multi_render_xray(m_, &params_, out.getData()), "srt_multi_render_xray");
    }
    // device time of the last render: the slowest device's whole step (render + assembly wait + repair), in ms
    float stepMilliseconds()
    {
        float ms = 0.0f;
        check(srt_multi_last_timing(m_, nullptr, &ms), "srt_multi_last_timing");
        return ms;
    }

private:
    void check(int rc, const char* what) const
    {
        if (rc != SRT_OK) throw std::runtime_error(std::string(what) + ": " + srt_multi_last_error(m_));
    }
    srt_multi* m_ = nullptr;
    srt_params params_{};
    srt_camera cam_{};
};

// createBackground (src/main-omp.cxx:525-555): textured quad behind the scene, built from the scene bbox.
inline TriangleMesh createBackground(const Vec3& upper, const Vec3& lower, const ImageRGB& texture)
{
    const Vec3 range = upper - lower;
    const float x = upper[0] + range[0] * 0.1f;
    const float y0 = lower[1] - range[1] * 0.5f, y1 = upper[1] + range[1] * 0.5f;
    const float z0 = lower[2] - range[2] * 0.5f, z1 = upper[2] + range[2] * 0.5f;
    const std::vector<float> vertices = { x, y0, z0, x, y1, z0, x, y1, z1, x, y0, z1 };
    const std::vector<float> text_coords = { 0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 0, 0 };
    const std::vector<unsigned int> indices = { 0, 1, 2, 0, 2, 3 };
    TriangleMesh mesh(vertices, indices, text_coords);
    mesh.setTexture(texture);
    return mesh;
}

// Drop-in for the shaded drivers' renderLoop (src/main-omp.cxx:136-142, 588-797): closest hit, Phong shading,
// shadow rays and texturing on GPU 0. The image keeps its current colour where no triangle is hit, so it must
// be a uniform background (Image(width, height, r, g, b) as in main-omp.cxx:233).
inline void renderLoop(ImageRGB& anOutputImage, const std::vector<TriangleMesh>& aTriangleMeshSet,
                       const Vec3& aDetectorPosition, const Vec3& aRayOrigin, const Vec3& anUpVector,
                       const Vec3& aRightVector, const Light& aLight)
{
    const unsigned int W = anOutputImage.getWidth(), H = anOutputImage.getHeight();
    if (W == 0 || H == 0) throw std::runtime_error("renderLoop: empty output image");
    XRayRenderer r(0);
    r.setMeshes(aTriangleMeshSet);
    Vec3 upper, lower;
    getBBox(aTriangleMeshSet, upper, lower);                 // bbox of ALL meshes sets the pixel spacing (:594-607)
    RayTracerInfo info;
    info.detector_position = aDetectorPosition; info.origin = aRayOrigin; info.up = anUpVector; info.right = aRightVector;
    info.range = upper - lower;
    r.setCamera(info, W, H);
    srt_light light{};
    for (int k = 0; k < 3; ++k) { light.colour[k] = aLight.getColour()[k]; light.direction[k] = aLight.getDirection()[k]; light.position[k] = aLight.getPosition()[k]; }
    unsigned char bg[3];
    anOutputImage.getPixel(0, 0, bg[0], bg[1], bg[2]);
    detail::check(srt_render_shaded(r.handle(), &r.params(), &light, bg, anOutputImage.getData()), r.handle(), "srt_render_shaded");
}

// Drop-in for the serial driver's renderLoop (src/main.cxx:159-161, 626-743): the SORT-based X-ray image of the scene on
// GPU 0 -- per ray the intersections with mesh 0 are sorted and exit - entry pairs summed, 80 exp(-(0.3971 d / 10)).
// This is the function that produced out/dragon-128x128-serial.txt.
inline void renderLoop(Image& anOutputImage, const std::vector<TriangleMesh>& aTriangleMeshSet, RayTracerInfo& rayTracerInfo)
{
    XRayRenderer r(0);
    const detail::FlatMeshes f = detail::flatten(aTriangleMeshSet, /*mu0: main.cxx:739*/ 0.3971f, 0.3971f);
    detail::check(srt_set_meshes(r.handle(), (int)f.desc.size(), f.desc.data()), r.handle(), "srt_set_meshes");
    r.setCamera(rayTracerInfo, anOutputImage.getWidth(), anOutputImage.getHeight());
    detail::check(srt_render_xray_sorted(r.handle(), &r.params(), anOutputImage.getData()), r.handle(), "srt_render_xray_sorted");
}

// The flow of main-pthreads-lbuffer.cxx (render callback :733-813 over all pixels + repair / image fill :329-404): the
// L-buffer X-ray image (signed path-length sum, artefact flag, Beer-Lambert with mu = 0.1037 / 0.3971, repair).
inline void renderLoopLBuffer(Image& anOutputImage, const std::vector<TriangleMesh>& aTriangleMeshSet, RayTracerInfo& rayTracerInfo)
{
    XRayRenderer r(0);
    r.setMeshes(aTriangleMeshSet);
    r.setCamera(rayTracerInfo, anOutputImage.getWidth(), anOutputImage.getHeight());
    r.render(anOutputImage);
}

} // namespace srt
