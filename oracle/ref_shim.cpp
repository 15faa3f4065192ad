/*
 * ref_shim.cpp -- C interface around the reference's OWN classes.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/srt_oracle.c header).  This file is
 * compiled together with /root/reference/src/{Ray,Triangle,TriangleMesh}.cxx and
 * the reference headers, unmodified, into oracle/_ref/libsrt_ref.so (recipe:
 * oracle/Makefile).  All arithmetic on the hot path below is executed by the
 * reference's code: Vec3 operators, Triangle::computeNormal, TriangleMesh::
 * setGeometry / computeBoundingBox, Ray::Ray, Ray::intersect.  What is restated
 * here, because the reference drivers need Assimp/glm/libjpeg (absent):
 *   - the driver bodies (camera set-up, render callbacks, repair pass), written
 *     with the reference's types so that float/double promotion is decided by
 *     the reference's operator overloads, not by this file;
 *   - the four non-JPEG members of Image that src/Image.cxx would provide
 *     (src/Image.cxx:37-81, 210-263), needed to link TriangleMesh.
 * Purpose: validate oracle/srt_oracle.c bit-for-bit and serve as the
 * "reference" CPU baseline (bench.py --impl reference).
 */
#include <cmath>
#include <cstring>
#include <cstdint>
#include <fstream>
#include <limits>
#include <stdexcept>
#include <vector>
#include <algorithm>
#include <pthread.h>

#include "Vec3.h"
#include "Ray.h"
#include "Triangle.h"
#include "TriangleMesh.h"
#include "Image.h"
#include "Material.h"
#include "Light.h"

/* ---- Image members normally supplied by src/Image.cxx -------------------- */
Image::Image(const Image& o) : m_p_pixel_data(0), m_width(o.m_width), m_height(o.m_height)
{
    if (m_width && m_height && o.m_p_pixel_data) {
        m_p_pixel_data = new float[m_width * m_height];
        std::memcpy(m_p_pixel_data, o.m_p_pixel_data, sizeof(float) * m_width * m_height);
    }
}
Image& Image::operator=(const Image& o)
{
    destroy();
    if (o.m_width && o.m_height && o.m_p_pixel_data) {
        setSize(o.m_width, o.m_height);
        std::memcpy(m_p_pixel_data, o.m_p_pixel_data, sizeof(float) * o.m_width * o.m_height);
    }
    m_width = o.m_width; m_height = o.m_height;
    return *this;
}
void Image::setSize(unsigned int w, unsigned int h, float v)
{
    if (w != getWidth() || h != getHeight()) {
        destroy();
        m_width = w; m_height = h;
        if (m_width && m_height) {
            m_p_pixel_data = new float[m_width * m_height];
            for (unsigned int i = 0; i < m_width * m_height; ++i) m_p_pixel_data[i] = v;
        }
    }
}
void Image::saveTextFile(const std::string& name)
{
    std::ofstream f(name);
    if (!f.is_open()) throw std::runtime_error("Cannot create the file " + name);
    for (unsigned int r = 0; r < m_height; ++r) {
        for (unsigned int c = 0; c < m_width; ++c) {
            f << m_p_pixel_data[r * m_width + c];
            if (c + 1 < m_width) f << "\t";
        }
        if (r + 1 < m_height) f << std::endl;
    }
}
void Image::saveJPEGFile(const char*, float, float) { throw std::runtime_error("no libjpeg in the oracle build"); }
void Image::saveTGAFile(const char*, float, float) { throw std::runtime_error("not built"); }

namespace {

struct RayTracerInfo {            /* main-pthreads-lbuffer.cxx:125-135 (Light omitted) */
    Vec3 detector_position, origin, up, right, upper_bbox_corner, lower_bbox_corner, range;
};

std::vector<TriangleMesh> build_meshes(const float* tri, const uint32_t* ntris, int n_meshes)
{
    std::vector<TriangleMesh> meshes;
    size_t off = 0;
    for (int m = 0; m < n_meshes; ++m) {
        std::vector<float> v(tri + off, tri + off + 9 * (size_t)ntris[m]);
        TriangleMesh mesh;
        mesh.setGeometry(v);                 /* TriangleMesh.cxx:80-100 */
        meshes.push_back(mesh);
        off += 9 * (size_t)ntris[m];
    }
    return meshes;
}

/* main-pthreads-lbuffer.cxx:585-608 */
void getBBox(const std::vector<TriangleMesh>& set, Vec3& hi, Vec3& lo)
{
    float inf = std::numeric_limits<float>::infinity();
    lo = Vec3(inf, inf, inf); hi = Vec3(-inf, -inf, -inf);
    for (std::vector<TriangleMesh>::const_iterator it = set.begin(); it != set.end(); ++it) {
        Vec3 l = it->getLowerBBoxCorner(), h = it->getUpperBBoxCorner();
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
    }
}

/* main-pthreads-lbuffer.cxx:611-667 -- expressions kept in the reference's
 * form so that its Vec3 overloads pick float or double */
RayTracerInfo initialiseRayTracing(const Vec3& hi, const Vec3& lo)
{
    Vec3 range = hi - lo;
    Vec3 bbox_centre = lo + range / 2.0;
    float diagonal = range.getLength();
    Vec3 up(0.0, 0.0, -1.0);
    Vec3 origin(bbox_centre - Vec3(diagonal * 1, 0, 0));
    Vec3 detector_position(bbox_centre + Vec3(diagonal * 0.6, 0, 0));
    Vec3 direction((detector_position - origin));
    direction.normalize();
    direction.normalise();
    Vec3 right(direction.crossProduct(up));
    RayTracerInfo info = { detector_position, origin, up, right, hi, lo, range };
    return info;
}

template <typename T> int signum(T val) { return (T(0) < val) - (val < T(0)); }

struct Job {
    const std::vector<TriangleMesh>* meshes; const RayTracerInfo* info;
    unsigned int W, H; uint64_t start_pixel, end_pixel;
    float* L; float* dist; int32_t* ssum; uint32_t* hits;
};

/* main-pthreads-lbuffer.cxx:733-813 */
void* renderLoopCallBack(void* p)
{
    Job* d = static_cast<Job*>(p);
    const RayTracerInfo& info = *d->info;
    const std::vector<TriangleMesh>& set = *d->meshes;
    unsigned int W = d->W, H = d->H;
    float res1 = info.range[2] / W;
    float res2 = info.range[1] / H;
    float pixel_spacing[] = { 2 * std::max(res1, res2), 2 * std::max(res1, res2) };
    for (uint64_t pixel = d->start_pixel; pixel <= d->end_pixel; ++pixel) {
        unsigned int row = pixel / W;
        unsigned int col = pixel % W;
        float v_offset = pixel_spacing[1] * (0.5 + row - H / 2.0);
        float u_offset = pixel_spacing[0] * (0.5 + col - W / 2.0);
        Vec3 direction = info.detector_position + info.up * v_offset + info.right * u_offset - info.origin;
        direction.normalise();
        Ray ray(info.origin, direction);
        for (std::vector<TriangleMesh>::const_iterator mesh = set.begin(); mesh != set.end(); ++mesh) {
            if (d->L[row * (uint64_t)W + col] == -1) break;
            if (mesh->intersectBBox(ray)) {
                float distance = 0.0f; int signSum = 0; uint32_t nh = 0;
                for (unsigned int k = 0; k < mesh->getNumberOfTriangles(); ++k) {
                    const Triangle& triangle = mesh->getTriangle(k);
                    float t;
                    bool intersect = ray.intersect(triangle, t);
                    if (intersect && &*mesh == &set[0] && t > 0.0000001) {
                        int sign = signum(direction.dotProduct(triangle.getNormal()));
                        distance += (sign * t);
                        signSum += sign;
                        ++nh;
                    }
                }
                float attenuationCoefficient = 0;
                if (&(*mesh) == &(set[0])) attenuationCoefficient = 0.1037f; else attenuationCoefficient = 0.3971f;
                float& Lp = d->L[row * (uint64_t)W + col];
                if (signSum != 0) Lp = -1;
                else Lp = Lp * exp(-(attenuationCoefficient * (distance * 0.1)));
                if (&*mesh == &set[0]) {
                    if (d->dist) d->dist[pixel] = distance;
                    if (d->ssum) d->ssum[pixel] = signSum;
                    if (d->hits) d->hits[pixel] = nh;
                }
            }
        }
    }
    return 0;
}

} // namespace

extern "C" {

struct ref_camera {
    float detector[3], origin[3], up[3], right[3], lower[3], upper[3], range[3];
    float light_position[3], light_direction[3];
};

static void put(float* o, const Vec3& v) { o[0] = v.getX(); o[1] = v.getY(); o[2] = v.getZ(); }
static Vec3 get(const float* p) { return Vec3(p[0], p[1], p[2]); }

void ref_scene_camera(const float* tri, const uint32_t* ntris, int n_meshes, ref_camera* cam)
{
    std::vector<TriangleMesh> meshes = build_meshes(tri, ntris, n_meshes);
    Vec3 lo, hi;
    getBBox(meshes, hi, lo);
    RayTracerInfo info = initialiseRayTracing(hi, lo);
    put(cam->detector, info.detector_position); put(cam->origin, info.origin);
    put(cam->up, info.up); put(cam->right, info.right);
    put(cam->lower, lo); put(cam->upper, hi); put(cam->range, info.range);
    put(cam->light_position, info.origin);
    Vec3 ld = (lo + info.range / 2.0) - info.origin; ld.normalise();
    put(cam->light_direction, ld);
}

int ref_intersect(const float o[3], const float d[3], const float tri9[9], float* t)
{
    /* NB: Ray::Ray renormalises d (Ray.inl:74-85); callers compare against
     * oracle_intersect(o, oracle_ray_direction(d), ...) */
    Ray ray(get(o), get(d));
    Triangle tr(get(tri9), get(tri9 + 3), get(tri9 + 6));
    float tt = 0;
    bool hit = ray.intersect(tr, tt);
    *t = tt;
    return hit ? 1 : 0;
}

/* Ray::Ray (Ray.inl:74-85): returns the stored direction */
void ref_ray_direction(const float d[3], float out[3])
{
    Ray ray(Vec3(0, 0, 0), get(d));
    put(out, ray.getDirection());
}

void ref_triangle_normal(const float tri9[9], float n[3])
{
    Triangle tr(get(tri9), get(tri9 + 3), get(tri9 + 6));
    put(n, tr.getNormal());
}

float ref_triangle_area(const float tri9[9])
{
    Triangle tr(get(tri9), get(tri9 + 3), get(tri9 + 6));
    return tr.getArea();
}

int ref_render_lbuffer(const float* tri, const uint32_t* ntris, int n_meshes, const ref_camera* cam,
                       uint32_t W, uint32_t H, uint64_t pixel_begin, uint64_t pixel_end, int n_threads,
                       float* L, float* dist, int32_t* ssum, uint32_t* hits)
{
    std::vector<TriangleMesh> meshes = build_meshes(tri, ntris, n_meshes);
    RayTracerInfo info = { get(cam->detector), get(cam->origin), get(cam->up), get(cam->right),
                           get(cam->upper), get(cam->lower), get(cam->range) };
    if (n_threads < 1) n_threads = 1;
    for (uint64_t p = pixel_begin; p <= pixel_end; ++p) L[p] = 80.000f;
    std::vector<Job> jobs(n_threads);
    std::vector<pthread_t> tid(n_threads);
    std::vector<char> created(n_threads, 0);
    uint64_t total = pixel_end - pixel_begin + 1, per = total / n_threads, rem = total % n_threads;
    int64_t last = (int64_t)pixel_begin - 1;
    for (int i = 0; i < n_threads; ++i) {        /* main-pthreads-lbuffer.cxx:694-725 */
        int64_t start = ++last, end = last + (int64_t)per - 1;
        if (rem > 0) { end++; --rem; }
        last = end;
        if (end < start) continue;
        Job j = { &meshes, &info, W, H, (uint64_t)start, (uint64_t)end, L, dist, ssum, hits };
        jobs[i] = j;
        pthread_create(&tid[i], NULL, renderLoopCallBack, &jobs[i]);
        created[i] = 1;
    }
    for (int i = 0; i < n_threads; ++i) if (created[i]) pthread_join(tid[i], NULL);
    return 0;
}

/* src/main.cxx:626-743 (sort-based X-ray; produced out/dragon-128x128-serial.txt) */
int ref_render_sort(const float* tri, const uint32_t* ntris, int n_meshes, const ref_camera* cam,
                    uint32_t W, uint32_t H, uint64_t pixel_begin, uint64_t pixel_end, float* image)
{
    using namespace std;
    vector<TriangleMesh> set = build_meshes(tri, ntris, n_meshes);
    Vec3 det = get(cam->detector), origin = get(cam->origin), up = get(cam->up), right = get(cam->right);
    Vec3 range = get(cam->range);
    float res1 = range[2] / W;
    float res2 = range[1] / H;
    float pixel_spacing[] = { 2 * std::max(res1, res2), 2 * std::max(res1, res2) };
    for (uint64_t pixel = pixel_begin; pixel <= pixel_end; ++pixel) {
        unsigned int row = pixel / W;
        unsigned int col = pixel % W;
        float v_offset = pixel_spacing[1] * (0.5 + row - H / 2.0);
        float u_offset = pixel_spacing[0] * (0.5 + col - W / 2.0);
        Vec3 direction = det + up * v_offset + right * u_offset - origin;
        direction.normalise();
        Ray ray(origin, direction);
        vector<float> pts;
        bool any = false;
        for (vector<TriangleMesh>::const_iterator mesh = set.begin(); mesh != set.end(); ++mesh)
            for (unsigned int k = 0; k < mesh->getNumberOfTriangles(); ++k) {
                float t;
                bool intersect = ray.intersect(mesh->getTriangle(k), t);
                if (intersect && &*mesh == &set[0] && t > 0.0000001) { pts.push_back(t); any = true; }
            }
        float distance = 0;
        if (any && pts.size() % 2 == 0) {
            sort(pts.begin(), pts.end());
            for (size_t i = 0; i < pts.size(); i += 2) distance += pts[i + 1] - pts[i];
        }
        distance = distance * 0.1;
        float photonOut = 80.000f * exp(-(0.3971f * distance));
        image[pixel] = photonOut;
    }
    return 0;
}

struct ref_material { float ambient[3], diffuse[3], specular[3]; float shininess; };

namespace {
/* applyShading, src/main-omp.cxx:407-434 -- expressions kept in the reference's form */
Vec3 applyShading(const Light& aLight, const Material& aMaterial, const Vec3& aNormalVector, const Vec3& aPosition,
                  const Vec3& aViewPosition)
{
    Vec3 ambient, diffuse, specular;
    ambient = aLight.getColour() * aMaterial.getAmbient();
    Vec3 lightDir = (aLight.getPosition() - aPosition);
    lightDir.normalize();
    float diff = std::max(std::abs(aNormalVector.dotProduct(lightDir)), 0.0f);
    diffuse = aLight.getColour() * (diff * aMaterial.getDiffuse());
    Vec3 viewDir(aViewPosition - aPosition);
    viewDir.normalize();
    Vec3 reflectDir = reflect(-viewDir, aNormalVector);
    float spec = std::pow(std::max(dot(viewDir, reflectDir), 0.0f), aMaterial.getShininess());
    specular = aLight.getColour() * (spec * aMaterial.getSpecular());
    return ambient + diffuse + specular;
}
} // namespace

/* src/main-omp.cxx:588-797 with the reference's classes; textures are raw RGB (the old RGB Image class,
 * include/oldimage.h, does not build: it includes the wrong .inl). Single-threaded per call. */
int ref_render_shaded(const float* tri, const float* texcoords, const uint32_t* ntris, int n_meshes,
                      const ref_material* mats, const uint8_t* const* tex, const uint32_t* tex_w, const uint32_t* tex_h,
                      const ref_camera* cam, uint32_t W, uint32_t H, uint64_t pixel_begin, uint64_t pixel_end, uint8_t* rgb)
{
    using namespace std;
    vector<TriangleMesh> set;
    size_t off = 0;
    for (int m = 0; m < n_meshes; ++m) {
        vector<float> v(tri + off, tri + off + 9 * (size_t)ntris[m]);
        TriangleMesh mesh;
        if (texcoords) { vector<float> tc(texcoords + off, texcoords + off + 9 * (size_t)ntris[m]); mesh.setGeometry(v, tc); }
        else mesh.setGeometry(v);
        mesh.setMaterial(Material(get(mats[m].ambient), get(mats[m].diffuse), get(mats[m].specular), mats[m].shininess));
        set.push_back(mesh);
        off += 9 * (size_t)ntris[m];
    }
    Vec3 aDetectorPosition = get(cam->detector), aRayOrigin = get(cam->origin), anUpVector = get(cam->up), aRightVector = get(cam->right);
    Vec3 light_position = aRayOrigin + anUpVector * 100.0;
    Light aLight(Vec3(1, 1, 1), Vec3(1, 0, 0), light_position);
    Vec3 upper_bbox_corner, lower_bbox_corner;
    getBBox(set, upper_bbox_corner, lower_bbox_corner);
    Vec3 range = upper_bbox_corner - lower_bbox_corner;
    float res1 = range[2] / W;
    float res2 = range[1] / H;
    float pixel_spacing[] = { 2 * std::max(res1, res2), 2 * std::max(res1, res2) };
    float inf = std::numeric_limits<float>::infinity();
    for (uint64_t pixel = pixel_begin; pixel < pixel_end; ++pixel) {
        int row = pixel / W, col = pixel % W;
        float v_offset = pixel_spacing[1] * (0.5 + row - H / 2.0);
        float u_offset = pixel_spacing[0] * (0.5 + col - W / 2.0);
        Vec3 direction = aDetectorPosition + anUpVector * v_offset + aRightVector * u_offset - aRayOrigin;
        direction.normalise();
        Ray ray(aRayOrigin, direction);
        const TriangleMesh* p_intersected_object = 0;
        const Triangle* p_intersected_triangle = 0;
        int hit_mesh = -1;
        float zb = inf;
        for (size_t m = 0; m < set.size(); ++m)
            for (unsigned int k = 0; k < set[m].getNumberOfTriangles(); ++k) {
                const Triangle& triangle = set[m].getTriangle(k);
                float t;
                if (ray.intersect(triangle, t))
                    if (zb > t) { zb = t; p_intersected_object = &set[m]; p_intersected_triangle = &triangle; hit_mesh = (int)m; }
            }
        if (p_intersected_object && p_intersected_triangle) {
            float t = zb;
            Vec3 point_hit = ray.getOrigin() + t * ray.getDirection();
            Material material = p_intersected_object->getMaterial();
            Vec3 colour = applyShading(aLight, material, p_intersected_triangle->getNormal(), point_hit, ray.getOrigin());
            unsigned char r = 0, g = 0, b = 0;
            Vec3 shadow_ray_direction = aLight.getPosition() - point_hit;
            shadow_ray_direction.normalise();
            Ray shadow_ray(point_hit, shadow_ray_direction);
            bool is_point_in_shadow = false;
            for (size_t m = 0; m < set.size(); ++m)
                for (unsigned int k = 0; k < set[m].getNumberOfTriangles(); ++k) {
                    const Triangle& triangle = set[m].getTriangle(k);
                    if (&triangle != p_intersected_triangle) {
                        float ts;
                        bool intersection = shadow_ray.intersect(triangle, ts);
                        if (intersection && ts > 0.0000001) { is_point_in_shadow = true; break; }
                    }
                }
            if (is_point_in_shadow) { colour[0] *= 0.25; colour[1] *= 0.25; colour[2] *= 0.25; }
            const uint8_t* texture = tex ? tex[hit_mesh] : 0;
            unsigned int tw = texture ? tex_w[hit_mesh] : 0, th = texture ? tex_h[hit_mesh] : 0;
            if (tw * th) {
                Vec3 P = aRayOrigin + t * direction;
                Vec3 A = p_intersected_triangle->getP1(), B = p_intersected_triangle->getP2(), C = p_intersected_triangle->getP3();
                Triangle ABC(A, B, C), ABP(A, B, P), BCP(B, C, P), CAP(C, A, P);
                float area_ABC = ABC.getArea();
                float u = CAP.getArea() / area_ABC;
                float v = ABP.getArea() / area_ABC;
                float w = BCP.getArea() / area_ABC;
                Vec3 texel_coord(w * p_intersected_triangle->getTextCoord1() + u * p_intersected_triangle->getTextCoord2() +
                                 v * p_intersected_triangle->getTextCoord3());
                unsigned int ti = texel_coord[0] * (tw - 1), tj = texel_coord[1] * (th - 1);
                if (ti >= tw) ti = tw - 1;
                if (tj >= th) tj = th - 1;
                unsigned char texel_r = texture[(tj * (size_t)tw + ti) * 3], texel_g = texture[(tj * (size_t)tw + ti) * 3 + 1],
                              texel_b = texture[(tj * (size_t)tw + ti) * 3 + 2];
                colour[0] *= texel_r; colour[1] *= texel_g; colour[2] *= texel_b;
                if (colour[0] < 0) r = 0; else if (colour[0] > 255) r = 255; else r = int(colour[0]);
                if (colour[1] < 0) g = 0; else if (colour[1] > 255) g = 255; else g = int(colour[1]);
                if (colour[2] < 0) b = 0; else if (colour[2] > 255) b = 255; else b = int(colour[2]);
            } else {
                if (255.0 * colour[0] < 0) r = 0; else if (255.0 * colour[0] > 255) r = 255; else r = int(255.0 * colour[0]);
                if (255.0 * colour[1] < 0) g = 0; else if (255.0 * colour[1] > 255) g = 255; else g = int(255.0 * colour[1]);
                if (255.0 * colour[2] < 0) b = 0; else if (255.0 * colour[2] > 255) b = 255; else b = int(255.0 * colour[2]);
            }
            rgb[pixel * 3] = r; rgb[pixel * 3 + 1] = g; rgb[pixel * 3 + 2] = b;
        }
    }
    return 0;
}

int ref_save_text(const char* path, const float* image, uint32_t W, uint32_t H)
{
    try {
        Image img(W, H, 0.0f);
        for (uint32_t r = 0; r < H; ++r)
            for (uint32_t c = 0; c < W; ++c) img.setPixel(c, r, image[(uint64_t)r * W + c]);
        img.saveTextFile(std::string(path));
    } catch (...) { return -1; }
    return 0;
}

} // extern "C"
