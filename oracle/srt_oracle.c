/*
 * srt_oracle.c -- CPU restatement of the SimpleRayTracing per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: a plain-C
 * restatement of the reference's algorithm, one function per reference
 * function, each citing the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (simpleraytracing_b200/) never links, imports or
 * calls anything in oracle/.
 *
 * Parity pinning: oracle_render_sort() reproduces the reference's only golden
 * numeric file (out/dragon-128x128-serial.txt, fixture
 * tests/golden/dragon_128_serial.npy) and every function here is cross-checked
 * bit-for-bit against oracle/_ref (the reference's own Ray.cxx / Triangle.cxx /
 * TriangleMesh.cxx compiled unmodified) by tests/test_oracle.py.
 *
 * Arithmetic discipline: every float operation of the reference is a
 * separately rounded IEEE binary32 operation (x86-64 SSE, no FMA contraction,
 * FLT_EVAL_METHOD == 0).  Build with -ffp-contract=off and without
 * -ffast-math / -march=native (see oracle/Makefile).  Where the reference
 * mixes float and double (Vec3::operator*(double), "1.0 / det", "0.5 + row")
 * the promotion is written out explicitly.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FMA__)
#error "build the oracle without -mfma / -march=native: contraction changes the bits"
#endif

typedef struct { float x, y, z; } v3;

/* ---- Vec3 (include/Vec3.inl) ------------------------------------------- */
static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
/* Vec3.inl:162-168 */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
/* Vec3.inl:172-178 */
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
/* Vec3.inl:205-209  operator*(float) */
static inline v3 v3_mulf(v3 a, float t) { return v3_make(a.x * t, a.y * t, a.z * t); }
/* Vec3.inl:213-217  operator/(float) */
static inline v3 v3_divf(v3 a, float t) { return v3_make(a.x / t, a.y / t, a.z / t); }
/* Vec3.inl:221-225  operator*(double): multiply in double, narrow in the Vec3 ctor */
static inline v3 v3_muld(v3 a, double t)
{ return v3_make((float)((double)a.x * t), (float)((double)a.y * t), (float)((double)a.z * t)); }
/* Vec3.inl:229-233  operator/(double) */
static inline v3 v3_divd(v3 a, double t)
{ return v3_make((float)((double)a.x / t), (float)((double)a.y / t), (float)((double)a.z / t)); }
/* Vec3.inl:313-317  left-to-right x*x' + y*y' + z*z' */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* Vec3.inl:321-329 */
static inline v3 v3_cross(v3 a, v3 b)
{ return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* Vec3.inl:461-465  std::sqrt(float) */
static inline float v3_len(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* Vec3.inl:469-476  three divides by the length (NOT a multiply by 1/len) */
static inline v3 v3_normalise(v3 a) { float l = v3_len(a); return v3_make(a.x / l, a.y / l, a.z / l); }

/* ---- Triangle (include/Triangle.inl) ------------------------------------ */
/* Triangle.inl:170-178 computeNormal */
static inline v3 tri_normal(v3 p1, v3 p2, v3 p3)
{
    v3 u = v3_sub(p2, p1);
    v3 v = v3_sub(p3, p1);
    return v3_normalise(v3_cross(u, v));
}
/* Triangle.inl:206-216 getArea: double pow/sqrt on float cross components, narrowed on return */
static inline float tri_area(v3 p1, v3 p2, v3 p3)
{
    v3 ab = v3_sub(p2, p1), ac = v3_sub(p3, p1);
    double a = pow((double)(ab.y * ac.z - ab.z * ac.y), 2);
    double b = pow((double)(ab.z * ac.x - ab.x * ac.z), 2);
    double c = pow((double)(ab.x * ac.y - ab.y * ac.x), 2);
    return (float)(0.5 * sqrt(a + b + c));
}

static inline v3 ld3(const float* p) { return v3_make(p[0], p[1], p[2]); }
static inline void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

/* ---- Ray (include/Ray.inl, src/Ray.cxx) ---------------------------------- */
/* Ray.inl:74-85  the constructor re-normalises; zero length leaves Vec3() = 0 */
static inline v3 ray_ctor_direction(v3 d)
{
    float length = v3_len(d);
    if (fpclassify(length) != FP_ZERO) return v3_divf(d, length);
    return v3_make(0.f, 0.f, 0.f);
}

/* Ray.cxx:72-125  Moller-Trumbore, two-sided, no epsilon; stage: 0 det==0,
 * 1 u-exit, 2 v-exit, 3 hit */
static inline int ray_intersect(v3 o, v3 d, v3 p1, v3 p2, v3 p3, float* t, int* stage)
{
    v3 edge1 = v3_sub(p2, p1);                    /* :86 */
    v3 edge2 = v3_sub(p3, p1);                    /* :87 */
    v3 pvec = v3_cross(d, edge2);                 /* :90 */
    float det = v3_dot(edge1, pvec);              /* :93 */
    if (fpclassify(det) == FP_ZERO) { *stage = 0; return 0; }   /* :94-97 */
    float inv_det = (float)(1.0 / (double)det);   /* :99  double divide, narrowed */
    v3 tvec = v3_sub(o, p1);                      /* :102 */
    float u = v3_dot(tvec, pvec) * inv_det;       /* :105 */
    if ((double)u < 0.0 || (double)u > 1.0) { *stage = 1; return 0; }   /* :106-109 */
    v3 qvec = v3_cross(tvec, edge1);              /* :112 */
    float v = v3_dot(d, qvec) * inv_det;          /* :115 */
    if ((double)v < 0.0 || (double)(u + v) > 1.0) { *stage = 2; return 0; } /* :116-119 */
    *t = v3_dot(edge2, qvec) * inv_det;           /* :122 */
    *stage = 3;
    return 1;
}

/* ---- camera -------------------------------------------------------------- */
typedef struct {
    float detector[3], origin[3], up[3], right[3];
    float lower[3], upper[3], range[3];
    float light_position[3], light_direction[3];
} oracle_camera;

/* TriangleMesh.cxx:192-228 + main-pthreads-lbuffer.cxx:585-608: per-mesh
 * min/max of all vertices, then min/max over meshes (== one global min/max). */
void oracle_bbox(const float* tri, uint64_t n_tris, float lo[3], float hi[3])
{
    float inf = INFINITY;
    for (int k = 0; k < 3; ++k) { lo[k] = inf; hi[k] = -inf; }
    for (uint64_t i = 0; i < n_tris * 3; ++i)
        for (int k = 0; k < 3; ++k) {
            float c = tri[i * 3 + k];
            /* std::min(a,b) = (b<a)?b:a ; std::max(a,b) = (a<b)?b:a */
            lo[k] = (c < lo[k]) ? c : lo[k];
            hi[k] = (hi[k] < c) ? c : hi[k];
        }
}

/* main-pthreads-lbuffer.cxx:611-667 (== main.cxx:566-622) initialiseRayTracing */
void oracle_init_camera(const float lo_[3], const float hi_[3], oracle_camera* cam)
{
    v3 lo = ld3(lo_), hi = ld3(hi_);
    v3 range = v3_sub(hi, lo);                              /* :620 */
    v3 centre = v3_add(lo, v3_divd(range, 2.0));            /* :621 */
    float diagonal = v3_len(range);                         /* :631 */
    v3 up = v3_make(0.0f, 0.0f, -1.0f);                     /* :633 */
    v3 origin = v3_sub(centre, v3_make(diagonal * 1, 0, 0));                      /* :635 */
    v3 detector = v3_add(centre, v3_make((float)((double)diagonal * 0.6), 0, 0)); /* :636 */
    v3 direction = v3_normalise(v3_sub(detector, origin));  /* :638-639 */
    v3 light_position = origin;                             /* :643 */
    v3 light_direction = v3_normalise(v3_sub(centre, light_position));   /* :644-645 */
    direction = v3_normalise(direction);                    /* :648 */
    v3 right = v3_cross(direction, up);                     /* :649 */
    st3(cam->detector, detector); st3(cam->origin, origin); st3(cam->up, up); st3(cam->right, right);
    st3(cam->lower, lo); st3(cam->upper, hi); st3(cam->range, range);
    st3(cam->light_position, light_position); st3(cam->light_direction, light_direction);
}

/* main-pthreads-lbuffer.cxx:739-741 */
float oracle_pixel_spacing(const oracle_camera* cam, uint32_t W, uint32_t H)
{
    float res1 = cam->range[2] / (float)W;
    float res2 = cam->range[1] / (float)H;
    float m = (res1 < res2) ? res2 : res1;      /* std::max */
    return 2 * m;
}

/* main-pthreads-lbuffer.cxx:753-762: dirA = local `direction` after normalise()
 * (used by the sign test, :791); dirB = Ray::m_direction (used by intersect). */
void oracle_make_ray(const oracle_camera* cam, uint32_t W, uint32_t H, uint32_t row, uint32_t col,
                     float dirA_[3], float dirB_[3])
{
    float spacing = oracle_pixel_spacing(cam, W, H);
    float v_offset = (float)((double)spacing * (0.5 + (double)row - (double)H / 2.0));   /* :756 */
    float u_offset = (float)((double)spacing * (0.5 + (double)col - (double)W / 2.0));   /* :757 */
    v3 d = v3_sub(v3_add(v3_add(ld3(cam->detector), v3_mulf(ld3(cam->up), v_offset)),
                         v3_mulf(ld3(cam->right), u_offset)), ld3(cam->origin));         /* :760 */
    v3 dirA = v3_normalise(d);                                                           /* :761 */
    v3 dirB = ray_ctor_direction(dirA);                                                  /* :762 */
    st3(dirA_, dirA); st3(dirB_, dirB);
}

/* single test, exposed for the unit tests (edge cases of Ray.cxx:72-125) */
int oracle_intersect(const float o[3], const float d[3], const float tri9[9], float* t, int* stage)
{
    float tt = 0.f; int st = -1;
    int hit = ray_intersect(ld3(o), ld3(d), ld3(tri9), ld3(tri9 + 3), ld3(tri9 + 6), &tt, &st);
    if (t) *t = tt;
    if (stage) *stage = st;
    return hit;
}

/* Ray::Ray direction handling (Ray.inl:74-85) */
void oracle_ray_direction(const float d[3], float out[3]) { st3(out, ray_ctor_direction(ld3(d))); }

void oracle_triangle_normal(const float tri9[9], float n[3])
{ st3(n, tri_normal(ld3(tri9), ld3(tri9 + 3), ld3(tri9 + 6))); }

float oracle_triangle_area(const float tri9[9])
{ return tri_area(ld3(tri9), ld3(tri9 + 3), ld3(tri9 + 6)); }

/* template signum, main-pthreads-lbuffer.cxx:729-731 */
static inline int signum_f(float v) { return (0.0f < v) - (v < 0.0f); }

/* ---- L-buffer render (north-star path) ----------------------------------- */
typedef struct {
    const float* tri; const uint32_t* mesh_ntris; int n_meshes;
    const float* normals;                 /* 3 per triangle, Triangle::getNormal() */
    const oracle_camera* cam; uint32_t W, H;
    uint64_t start_pixel, end_pixel;      /* inclusive, as PThreadData (:137-166) */
    float* L; float* dist; int32_t* ssum; uint32_t* hits;
    uint64_t stage[4];
} lbuf_job;

/* main-pthreads-lbuffer.cxx:733-813 renderLoopCallBack */
static void* lbuffer_callback(void* arg)
{
    lbuf_job* j = (lbuf_job*)arg;
    const oracle_camera* cam = j->cam;
    const uint32_t W = j->W, H = j->H;
    float spacing = oracle_pixel_spacing(cam, W, H);                       /* :739-741 */
    v3 detector = ld3(cam->detector), origin = ld3(cam->origin), up = ld3(cam->up), right = ld3(cam->right);

    for (uint64_t pixel = j->start_pixel; pixel <= j->end_pixel; ++pixel) {   /* :750 */
        uint32_t row = (uint32_t)(pixel / W), col = (uint32_t)(pixel % W); /* :753-754 */
        float v_offset = (float)((double)spacing * (0.5 + (double)row - (double)H / 2.0));
        float u_offset = (float)((double)spacing * (0.5 + (double)col - (double)W / 2.0));
        v3 direction = v3_sub(v3_add(v3_add(detector, v3_mulf(up, v_offset)), v3_mulf(right, u_offset)), origin);
        direction = v3_normalise(direction);                               /* :761 */
        v3 rdir = ray_ctor_direction(direction);                           /* :762 */

        const float* tri = j->tri; const float* nrm = j->normals;
        for (int m = 0; m < j->n_meshes; ++m) {                            /* :768 */
            uint32_t nt = j->mesh_ntris[m];
            if (j->L[pixel] == -1) break;                                  /* :772 */
            float distance = 0.0f; int signSum = 0; uint32_t nh = 0;       /* :778-779 */
            for (uint32_t k = 0; k < nt; ++k) {                            /* :780 */
                float t; int st;
                int hit = ray_intersect(origin, rdir, ld3(tri + 9 * (size_t)k), ld3(tri + 9 * (size_t)k + 3),
                                        ld3(tri + 9 * (size_t)k + 6), &t, &st);   /* :785 */
                j->stage[st]++;
                if (hit && m == 0 && (double)t > 0.0000001) {              /* :788 */
                    int sign = signum_f(v3_dot(direction, ld3(nrm + 3 * (size_t)k)));   /* :791 */
                    distance += ((float)sign * t);                         /* :792 */
                    signSum += sign;                                       /* :793 */
                    ++nh;
                }
            }
            float mu = (m == 0) ? 0.1037f : 0.3971f;                       /* :798-803 */
            if (signSum != 0) j->L[pixel] = -1;                            /* :805-806 */
            else j->L[pixel] = (float)((double)j->L[pixel] * exp(-((double)mu * ((double)distance * 0.1)))); /* :808 */
            if (m == 0) {
                if (j->dist) j->dist[pixel] = distance;
                if (j->ssum) j->ssum[pixel] = signSum;
                if (j->hits) j->hits[pixel] = nh;
            }
            tri += 9 * (size_t)nt; nrm += 3 * (size_t)nt;
        }
    }
    return 0;   /* the reference omits the return (UB); SURVEY appendix B */
}

/*
 * Render pixels [pixel_begin, pixel_end] (inclusive) of a W x H frame with
 * n_threads pthreads using the reference's contiguous partition
 * (pthreadWorkLoadAllocation, main-pthreads-lbuffer.cxx:694-725).
 * L (W*H floats) is initialised to 80.0f for the rendered range (:314).
 * dist / ssum / hits (optional, W*H) receive mesh-0 path length, signSum and
 * accepted-hit count. stage4 (optional) receives the Ray::intersect exit
 * histogram {det==0, u-exit, v-exit, hit}. Returns 0.
 */
int oracle_render_lbuffer(const float* tri, const uint32_t* mesh_ntris, int n_meshes,
                          const oracle_camera* cam, uint32_t W, uint32_t H,
                          uint64_t pixel_begin, uint64_t pixel_end, int n_threads,
                          float* L, float* dist, int32_t* ssum, uint32_t* hits, uint64_t* stage4)
{
    if (n_threads < 1) n_threads = 1;
    uint64_t total_tris = 0;
    for (int m = 0; m < n_meshes; ++m) total_tris += mesh_ntris[m];
    float* normals = (float*)malloc(sizeof(float) * 3 * (total_tris ? total_tris : 1));
    for (uint64_t k = 0; k < total_tris; ++k) oracle_triangle_normal(tri + 9 * k, normals + 3 * k);
    for (uint64_t p = pixel_begin; p <= pixel_end; ++p) L[p] = 80.000f;    /* :314 */

    lbuf_job* jobs = (lbuf_job*)calloc((size_t)n_threads, sizeof(lbuf_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    char* created = (char*)calloc((size_t)n_threads, 1);
    uint64_t total = pixel_end - pixel_begin + 1;
    uint64_t per = total / (uint64_t)n_threads, rem = total % (uint64_t)n_threads;   /* :701-703 */
    int64_t last = (int64_t)pixel_begin - 1;
    for (int i = 0; i < n_threads; ++i) {
        lbuf_job* j = &jobs[i];
        j->tri = tri; j->mesh_ntris = mesh_ntris; j->n_meshes = n_meshes; j->normals = normals;
        j->cam = cam; j->W = W; j->H = H; j->L = L; j->dist = dist; j->ssum = ssum; j->hits = hits;
        int64_t start = ++last;                                            /* :710 */
        int64_t end = last + (int64_t)per - 1;                             /* :711 */
        if (rem > 0) { end++; --rem; }                                     /* :714-717 */
        last = end;                                                        /* :719 */
        if (end < start) continue;      /* more threads than pixels: empty chunk */
        j->start_pixel = (uint64_t)start; j->end_pixel = (uint64_t)end;
        pthread_create(&tid[i], NULL, lbuffer_callback, j);                /* :317-320 */
        created[i] = 1;
    }
    uint64_t st[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < n_threads; ++i) {                                  /* :323-325 */
        if (!created[i]) continue;
        pthread_join(tid[i], NULL);
        for (int s = 0; s < 4; ++s) st[s] += jobs[i].stage[s];
    }
    free(created);
    if (stage4) for (int s = 0; s < 4; ++s) stage4[s] = st[s];
    free(jobs); free(tid); free(normals);
    return 0;
}

/* The same render for an arbitrary LIST of pixels (strided subsets of very large frames: SURVEY 8d evaluates the
 * 8192 x 8192 configuration on every 64th row and column).  Outputs are written at the pixels' own indices of the
 * full-frame arrays; triangle normals are computed once, the list is cut into n_threads contiguous pieces. */
typedef struct { lbuf_job job; const uint64_t* pixels; uint64_t first, last; } lbuf_list_job;

static void* lbuffer_list_callback(void* arg)
{
    lbuf_list_job* lj = (lbuf_list_job*)arg;
    for (uint64_t i = lj->first; i < lj->last; ++i) {
        lbuf_job one = lj->job;
        one.start_pixel = one.end_pixel = lj->pixels[i];
        lbuffer_callback(&one);
    }
    return 0;
}

int oracle_render_lbuffer_pixels(const float* tri, const uint32_t* mesh_ntris, int n_meshes,
                                 const oracle_camera* cam, uint32_t W, uint32_t H,
                                 const uint64_t* pixels, uint64_t n_pixels, int n_threads,
                                 float* L, float* dist, int32_t* ssum, uint32_t* hits)
{
    if (n_threads < 1) n_threads = 1;
    uint64_t total_tris = 0;
    for (int m = 0; m < n_meshes; ++m) total_tris += mesh_ntris[m];
    float* normals = (float*)malloc(sizeof(float) * 3 * (total_tris ? total_tris : 1));
    for (uint64_t k = 0; k < total_tris; ++k) oracle_triangle_normal(tri + 9 * k, normals + 3 * k);
    for (uint64_t i = 0; i < n_pixels; ++i) L[pixels[i]] = 80.000f;        /* :314 */
    lbuf_list_job* jobs = (lbuf_list_job*)calloc((size_t)n_threads, sizeof(lbuf_list_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        lbuf_job* j = &jobs[i].job;
        j->tri = tri; j->mesh_ntris = mesh_ntris; j->n_meshes = n_meshes; j->normals = normals;
        j->cam = cam; j->W = W; j->H = H; j->L = L; j->dist = dist; j->ssum = ssum; j->hits = hits;
        jobs[i].pixels = pixels;
        jobs[i].first = n_pixels * (uint64_t)i / (uint64_t)n_threads;
        jobs[i].last = n_pixels * (uint64_t)(i + 1) / (uint64_t)n_threads;
        pthread_create(&tid[i], NULL, lbuffer_list_callback, &jobs[i]);
    }
    for (int i = 0; i < n_threads; ++i) pthread_join(tid[i], NULL);
    free(jobs); free(tid); free(normals);
    return 0;
}

/* ---- artefact repair + image fill (main-pthreads-lbuffer.cxx:329-404) ----- */
/* All index arithmetic is unsigned 32-bit as in the reference (row/col are
 * `unsigned int`, so `col - i` and `row - i` wrap).  The reference's guard is
 * `index > size`, which lets index == size read one element past the end (UB);
 * the restatement treats index == size as "stop" (SURVEY appendix B). */
static inline int repair_probe(const float* L, uint32_t index, uint64_t size, float* value)
{
    if ((uint64_t)index >= size) return -1;        /* break (:344 etc.) */
    if (L[index] != -1) { *value = L[index]; return 1; }
    return 0;
}

void oracle_repair(const float* L, uint32_t W, uint32_t H, float* image)
{
    uint64_t size = (uint64_t)W * H;
    for (uint64_t pixel = 0; pixel < size; ++pixel) {
        uint32_t row = (uint32_t)(pixel / W), col = (uint32_t)(pixel % W);
        float photonOut = L[row * W + col];
        if (photonOut == -1) {
            const uint32_t maxDepth = 4;                                   /* :339 */
            float vals[4]; int n = 0;
            float value = 0; int r;
            for (uint32_t i = 1; i <= maxDepth; ++i) {                     /* +col, :343-353 */
                r = repair_probe(L, row * W + (col + i), size, &value); if (r) break; }
            if (value != 0) vals[n++] = value;
            value = 0;
            for (uint32_t i = 1; i <= maxDepth; ++i) {                     /* -row, :356-366 */
                r = repair_probe(L, (row - i) * W + col, size, &value); if (r) break; }
            if (value != 0) vals[n++] = value;
            value = 0;
            for (uint32_t i = 1; i <= maxDepth; ++i) {                     /* -col, :369-379 */
                r = repair_probe(L, row * W + (col - i), size, &value); if (r) break; }
            if (value != 0) vals[n++] = value;
            value = 0;
            for (uint32_t i = 1; i <= maxDepth; ++i) {                     /* +row, :382-392 */
                r = repair_probe(L, (row + i) * W + col, size, &value); if (r) break; }
            if (value != 0) vals[n++] = value;
            float averageSum = 0;                                          /* :394-397 */
            for (int k = 0; k < n; ++k) averageSum += vals[k];
            photonOut = averageSum / (float)n;                             /* :399 (0/0 -> NaN) */
        }
        image[(uint64_t)row * W + col] = photonOut;                        /* :403 Image::setPixel */
    }
}

/* ---- sort-based X-ray (src/main.cxx:626-743) -- produced the golden file --- */
static int cmp_float(const void* a, const void* b)
{ float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }

int oracle_render_sort(const float* tri, const uint32_t* mesh_ntris, int n_meshes,
                       const oracle_camera* cam, uint32_t W, uint32_t H,
                       uint64_t pixel_begin, uint64_t pixel_end, float* image, uint32_t* hits_out)
{
    float spacing = oracle_pixel_spacing(cam, W, H);                       /* main.cxx:639-641 */
    v3 detector = ld3(cam->detector), origin = ld3(cam->origin), up = ld3(cam->up), right = ld3(cam->right);
    size_t cap = 64; float* pts = (float*)malloc(cap * sizeof(float));
    for (uint64_t pixel = pixel_begin; pixel <= pixel_end; ++pixel) {      /* :649 */
        uint32_t row = (uint32_t)(pixel / W), col = (uint32_t)(pixel % W);
        float v_offset = (float)((double)spacing * (0.5 + (double)row - (double)H / 2.0));
        float u_offset = (float)((double)spacing * (0.5 + (double)col - (double)W / 2.0));
        v3 direction = v3_sub(v3_add(v3_add(detector, v3_mulf(up, v_offset)), v3_mulf(right, u_offset)), origin);
        direction = v3_normalise(direction);
        v3 rdir = ray_ctor_direction(direction);
        size_t n = 0;
        const float* t9 = tri;
        for (int m = 0; m < n_meshes; ++m) {                               /* :670 */
            uint32_t nt = mesh_ntris[m];
            for (uint32_t k = 0; k < nt; ++k) {                            /* :678 */
                float t; int st;
                int hit = ray_intersect(origin, rdir, ld3(t9 + 9 * (size_t)k), ld3(t9 + 9 * (size_t)k + 3),
                                        ld3(t9 + 9 * (size_t)k + 6), &t, &st);
                if (hit && m == 0 && (double)t > 0.0000001) {              /* :687 */
                    if (n == cap) { cap *= 2; pts = (float*)realloc(pts, cap * sizeof(float)); }
                    pts[n++] = t;
                }
            }
            t9 += 9 * (size_t)nt;
        }
        float distance = 0;                                                /* :700 */
        if (n > 0 && n % 2 == 0) {                                         /* :701-703 */
            qsort(pts, n, sizeof(float), cmp_float);                       /* std::sort; ties irrelevant */
            for (size_t i = 0; i < n; i += 2) distance += pts[i + 1] - pts[i];   /* :706-708 */
        }
        distance = (float)((double)distance * 0.1);                        /* :725 */
        /* :739  `exp` on a float argument with <cmath> + using namespace std
         * resolves to std::exp(float) == expf */
        float photonOut = 80.000f * expf(-(0.3971f * distance));
        image[pixel] = photonOut;                                          /* :741 */
        if (hits_out) hits_out[pixel] = (uint32_t)n;
    }
    free(pts);
    return 0;
}

/* ---- Image::saveTextFile (src/Image.cxx:210-235) -------------------------- */
/* tab separated, default ostream float formatting (== "%g", 6 significant
 * digits), std::endl between rows, none after the last. */
int oracle_save_text(const char* path, const float* image, uint32_t W, uint32_t H)
{
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    for (uint32_t row = 0; row < H; ++row) {
        for (uint32_t col = 0; col < W; ++col) {
            fprintf(f, "%g", (double)image[(uint64_t)row * W + col]);
            if (col + 1 < W) fputc('\t', f);
        }
        if (row + 1 < H) fputc('\n', f);
    }
    fclose(f);
    return 0;
}

/* ---- work partition (main-pthreads-lbuffer.cxx:694-725), for tests -------- */
void oracle_partition(uint64_t total, uint32_t n, uint64_t* start, uint64_t* end)
{
    int64_t last = -1; uint64_t per = total / n, rem = total % n;
    for (uint32_t i = 0; i < n; ++i) {
        start[i] = (uint64_t)(++last);
        int64_t e = last + (int64_t)per - 1;
        if (rem > 0) { e++; --rem; }
        end[i] = (uint64_t)e; last = e;
    }
}

/* ======================================================================== */
/* Shaded variant: closest hit, Phong, shadow ray, texture */
/* restated from src/main-omp.cxx (same loop in main-mpi*.cxx / pthreads-rows) */
/* ======================================================================== */
typedef struct {
    float ambient[3], diffuse[3], specular[3];
    float shininess;
} oracle_material;

typedef struct {
    const float* tri; const float* texcoords; const uint32_t* mesh_ntris; int n_meshes;
    const oracle_material* mats; const uint8_t* const* tex; const uint32_t* tex_w; const uint32_t* tex_h;
    const oracle_camera* cam; float light_pos[3]; float light_col[3];
    uint32_t W, H; float spacing; uint64_t p0, p1;
    uint8_t* rgb; float* zbuf; int32_t* tri_id; uint8_t* shadow;
    const float* normals; const uint64_t* mesh_first;
} shade_job;

static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }     /* Vec3.inl:293-297 */
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
/* reflect(I, N) = I - 2.0 * N.dotProduct(I) * N: the scalar is a double, the product narrows per component
 * (Vec3.inl:120-124 with operator*(double, Vec3) -> Vec3::operator*(double)) */
static inline v3 v3_reflect(v3 I, v3 N) { return v3_sub(I, v3_muld(N, 2.0 * (double)v3_dot(N, I))); }

/* applyShading, src/main-omp.cxx:407-434 */
static v3 apply_shading(v3 Lc, v3 Lpos, const oracle_material* m, v3 N, v3 P, v3 view_pos)
{
    v3 Ka = ld3(m->ambient), Kd = ld3(m->diffuse), Ks = ld3(m->specular);
    v3 ambient = v3_mul(Lc, Ka);                                              /* :416 */
    v3 lightDir = v3_normalise(v3_sub(Lpos, P));                              /* :419-420 */
    float diff = fabsf(v3_dot(N, lightDir)); if (!(diff > 0.0f)) diff = (diff < 0.0f) ? 0.0f : diff;   /* max(|.|, 0) */
    v3 diffuse = v3_mul(Lc, v3_mulf(Kd, diff));                               /* :422 */
    v3 viewDir = v3_normalise(v3_sub(view_pos, P));                           /* :425-426 */
    v3 reflectDir = v3_reflect(v3_neg(viewDir), N);                           /* :428 */
    float d = v3_dot(viewDir, reflectDir);
    float base = (0.0f < d) ? d : 0.0f;                                       /* std::max(d, 0.0f) = (a < b) ? b : a */
    float spec = powf(base, m->shininess);                                    /* std::pow(float, float) :429 */
    v3 specular = v3_mul(Lc, v3_mulf(Ks, spec));                              /* :430 */
    return v3_add(v3_add(ambient, diffuse), specular);                        /* :432 */
}

static inline uint8_t clamp_trunc_f(float c) { if (c < 0) return 0; if (c > 255) return 255; return (uint8_t)(int)c; }
static inline uint8_t clamp_trunc_d(double c) { if (c < 0) return 0; if (c > 255) return 255; return (uint8_t)(int)c; }

/* renderLoop, src/main-omp.cxx:588-797, for pixels [p0, p1) */
static void* shade_worker(void* arg)
{
    shade_job* j = (shade_job*)arg;
    const oracle_camera* cam = j->cam;
    const uint32_t W = j->W, H = j->H;
    v3 detector = ld3(cam->detector), origin = ld3(cam->origin), up = ld3(cam->up), right = ld3(cam->right);
    v3 Lc = ld3(j->light_col), Lpos = ld3(j->light_pos);
    uint64_t total = 0; for (int m = 0; m < j->n_meshes; ++m) total += j->mesh_ntris[m];
    for (uint64_t pixel = j->p0; pixel < j->p1; ++pixel) {
        uint32_t row = (uint32_t)(pixel / W), col = (uint32_t)(pixel % W);
        float v_offset = (float)((double)j->spacing * (0.5 + (double)row - (double)H / 2.0));     /* :620 */
        float u_offset = (float)((double)j->spacing * (0.5 + (double)col - (double)W / 2.0));     /* :621 */
        v3 direction = v3_sub(v3_add(v3_add(detector, v3_mulf(up, v_offset)), v3_mulf(right, u_offset)), origin);
        direction = v3_normalise(direction);                                                       /* :625 */
        v3 rdir = ray_ctor_direction(direction);                                                   /* :626 */
        float z = INFINITY; int64_t hit_tri = -1; int hit_mesh = -1;
        uint64_t g = 0;
        for (int m = 0; m < j->n_meshes; ++m)                                                      /* :632 */
            for (uint32_t k = 0; k < j->mesh_ntris[m]; ++k, ++g) {                                 /* :640 */
                float t; int st;
                if (ray_intersect(origin, rdir, ld3(j->tri + 9 * g), ld3(j->tri + 9 * g + 3), ld3(j->tri + 9 * g + 6), &t, &st))
                    if (z > t) { z = t; hit_tri = (int64_t)g; hit_mesh = m; }                      /* :656-662: strict >, no t > 0 guard */
            }
        if (j->zbuf) j->zbuf[pixel] = z;
        if (j->tri_id) j->tri_id[pixel] = (int32_t)hit_tri;
        if (hit_tri < 0) continue;                                                                 /* background colour stays */
        float t = z;
        v3 point_hit = v3_add(origin, v3_mulf(rdir, t));                                           /* :671 */
        v3 N = ld3(j->normals + 3 * hit_tri);
        v3 colour = apply_shading(Lc, Lpos, &j->mats[hit_mesh], N, point_hit, origin);             /* :673 */
        v3 sdir = v3_normalise(v3_sub(Lpos, point_hit));                                           /* :680-681 */
        v3 srd = ray_ctor_direction(sdir);                                                         /* :682 */
        int in_shadow = 0;
        g = 0;
        for (int m = 0; m < j->n_meshes; ++m) {                                                    /* :687 */
            uint64_t g0 = g;
            for (uint32_t k = 0; k < j->mesh_ntris[m]; ++k) {
                uint64_t gi = g0 + k;
                if ((int64_t)gi != hit_tri) {                                                      /* :698 */
                    float ts; int st;
                    if (ray_intersect(point_hit, srd, ld3(j->tri + 9 * gi), ld3(j->tri + 9 * gi + 3), ld3(j->tri + 9 * gi + 6), &ts, &st)
                        && (double)ts > 0.0000001) { in_shadow = 1; break; }                       /* :703-707: inner break only */
                }
            }
            g = g0 + j->mesh_ntris[m];
        }
        if (j->shadow) j->shadow[pixel] = (uint8_t)in_shadow;
        if (in_shadow) {                                                                           /* :715-720 */
            colour.x = (float)((double)colour.x * 0.25); colour.y = (float)((double)colour.y * 0.25);
            colour.z = (float)((double)colour.z * 0.25);
        }
        uint8_t r, gg, b;
        const uint8_t* tex = j->tex ? j->tex[hit_mesh] : NULL;
        uint32_t tw = tex ? j->tex_w[hit_mesh] : 0, th = tex ? j->tex_h[hit_mesh] : 0;
        if (tex && tw * th) {                                                                      /* :725 */
            v3 P = v3_add(origin, v3_mulf(direction, t));                                          /* :729: the pre-Ray direction */
            v3 A = ld3(j->tri + 9 * hit_tri), B = ld3(j->tri + 9 * hit_tri + 3), C = ld3(j->tri + 9 * hit_tri + 6);
            float area_ABC = tri_area(A, B, C);
            float u = tri_area(C, A, P) / area_ABC;                                                /* :742-745 */
            float v = tri_area(A, B, P) / area_ABC;
            float w = tri_area(B, C, P) / area_ABC;
            const float* tc = j->texcoords + 9 * hit_tri;
            v3 texel = v3_add(v3_add(v3_mulf(ld3(tc), w), v3_mulf(ld3(tc + 3), u)), v3_mulf(ld3(tc + 6), v));   /* :748 */
            uint32_t ti = (uint32_t)(texel.x * (float)(tw - 1)), tj = (uint32_t)(texel.y * (float)(th - 1));     /* :755-756 */
            if (ti >= tw) ti = tw - 1;      /* the reference would throw std::out_of_range here */
            if (tj >= th) tj = th - 1;
            const uint8_t* px = tex + ((size_t)tj * tw + ti) * 3;
            colour.x *= (float)px[0]; colour.y *= (float)px[1]; colour.z *= (float)px[2];          /* :758-760 */
            r = clamp_trunc_f(colour.x); gg = clamp_trunc_f(colour.y); b = clamp_trunc_f(colour.z);
        } else {                                                                                   /* :776-790 */
            r = clamp_trunc_d(255.0 * (double)colour.x); gg = clamp_trunc_d(255.0 * (double)colour.y);
            b = clamp_trunc_d(255.0 * (double)colour.z);
        }
        j->rgb[pixel * 3] = r; j->rgb[pixel * 3 + 1] = gg; j->rgb[pixel * 3 + 2] = b;              /* :793 */
    }
    return 0;
}

/*
 * Shaded render of pixels [pixel_begin, pixel_end).  cam = camera of the scene WITHOUT the background quad
 * (main-omp.cxx:220-241 run before createBackground is appended, :249); the pixel spacing is recomputed inside
 * from the bbox of ALL meshes (:594-607).  Light: white, at origin + up * 100.0 (:235).  rgb must be pre-filled
 * with the background colour.  zbuf / tri_id / shadow are optional per-pixel debug outputs.
 */
int oracle_render_shaded(const float* tri, const float* texcoords, const uint32_t* mesh_ntris, int n_meshes,
                         const oracle_material* mats, const uint8_t* const* tex, const uint32_t* tex_w, const uint32_t* tex_h,
                         const oracle_camera* cam, uint32_t W, uint32_t H, uint64_t pixel_begin, uint64_t pixel_end,
                         int n_threads, uint8_t* rgb, float* zbuf, int32_t* tri_id, uint8_t* shadow)
{
    if (n_threads < 1) n_threads = 1;
    uint64_t total = 0; for (int m = 0; m < n_meshes; ++m) total += mesh_ntris[m];
    float* normals = (float*)malloc(sizeof(float) * 3 * (total ? total : 1));
    for (uint64_t k = 0; k < total; ++k) oracle_triangle_normal(tri + 9 * k, normals + 3 * k);
    float lo[3], hi[3];
    oracle_bbox(tri, total, lo, hi);                                     /* getBBox over all meshes, :594 */
    float res1 = (hi[2] - lo[2]) / (float)W, res2 = (hi[1] - lo[1]) / (float)H;     /* :601-602 */
    float spacing = 2 * ((res1 < res2) ? res2 : res1);                   /* :603 */
    v3 lp = v3_add(ld3(cam->origin), v3_muld(ld3(cam->up), 100.0));      /* :235 */
    shade_job* jobs = (shade_job*)calloc((size_t)n_threads, sizeof(shade_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t n = pixel_end - pixel_begin;
    for (int i = 0; i < n_threads; ++i) {
        shade_job* j = &jobs[i];
        j->tri = tri; j->texcoords = texcoords; j->mesh_ntris = mesh_ntris; j->n_meshes = n_meshes; j->mats = mats;
        j->tex = tex; j->tex_w = tex_w; j->tex_h = tex_h; j->cam = cam; st3(j->light_pos, lp);
        j->light_col[0] = j->light_col[1] = j->light_col[2] = 1.0f;
        j->W = W; j->H = H; j->spacing = spacing; j->rgb = rgb; j->zbuf = zbuf; j->tri_id = tri_id; j->shadow = shadow;
        j->normals = normals;
        j->p0 = pixel_begin + n * (uint64_t)i / (uint64_t)n_threads;
        j->p1 = pixel_begin + n * (uint64_t)(i + 1) / (uint64_t)n_threads;
        pthread_create(&tid[i], NULL, shade_worker, j);
    }
    for (int i = 0; i < n_threads; ++i) pthread_join(tid[i], NULL);
    free(jobs); free(tid); free(normals);
    return 0;
}
