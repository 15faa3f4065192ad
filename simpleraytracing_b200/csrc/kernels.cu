// kernels.cu -- sm_100a kernels of the X-ray L-buffer path.
//
// Reference semantics implemented here (all citations under /root/reference):
//   ray generation          src/main-pthreads-lbuffer.cxx:753-762, include/Ray.inl:74-85
//   Ray::intersect          src/Ray.cxx:72-125
//   accumulation            src/main-pthreads-lbuffer.cxx:780-795
//   attenuation / flagging  src/main-pthreads-lbuffer.cxx:798-809
//   artefact repair         src/main-pthreads-lbuffer.cxx:329-404
//
// Arithmetic contract: everything that decides a hit or feeds an output value is
// computed with separately rounded IEEE binary32 operations in the reference's
// order (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fsqrt_rn are never
// contracted into FMAs and are unaffected by fast-math flags).  The FMA-contracted
// binary32 arithmetic of the filter (filtered kernel, bin pass, culled kernel) only ever
// REJECTS pairs that provably miss ("filter proof"); it never decides a hit
// and never produces a value.
//
// Kernels: lbuffer_exact_kernel (reference sequence for every pair, + audit of the filter),
// lbuffer_filtered_kernel (brute force: conservative gate per pair) + resolve_hits_kernel,
// bin_kernel + lbuffer_culled_kernel (the same gate applied to boxes of pixels first),
// shade_kernel, repair_kernel, deinterleave_kernel, prepack_kernel, offsets_kernel.
#include "kernels.h"

#include <algorithm>
#include <cfloat>
#include <cstring>

#include "prepack_math.h"

namespace srt {
namespace {

// ------------------------------------------------------------------ helpers
struct RayDirs { float ax, ay, az, bx, by, bz; };   // dirA: sign test; dirB: Ray::m_direction

__device__ __forceinline__ float dot_rn(float ax, float ay, float az, float bx, float by, float bz)
{   // Vec3::dotProduct, include/Vec3.inl:313-317: (x*x' + y*y') + z*z'
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

__device__ __forceinline__ RayDirs make_ray(const DevCamera& c, uint32_t row, uint32_t col)
{
    const float v = __ldg(c.v_off + row), u = __ldg(c.u_off + col);
    // direction = detector + up * v_offset + right * u_offset - origin          (lbuffer.cxx:760)
    const float x = __fsub_rn(__fadd_rn(__fadd_rn(c.dx, __fmul_rn(c.ux, v)), __fmul_rn(c.rx, u)), c.ox);
    const float y = __fsub_rn(__fadd_rn(__fadd_rn(c.dy, __fmul_rn(c.uy, v)), __fmul_rn(c.ry, u)), c.oy);
    const float z = __fsub_rn(__fadd_rn(__fadd_rn(c.dz, __fmul_rn(c.uz, v)), __fmul_rn(c.rz, u)), c.oz);
    // direction.normalise(): sqrt + three divides                                (Vec3.inl:461-476)
    const float len = __fsqrt_rn(dot_rn(x, y, z, x, y, z));
    RayDirs r;
    r.ax = __fdiv_rn(x, len); r.ay = __fdiv_rn(y, len); r.az = __fdiv_rn(z, len);
    // Ray::Ray re-normalises unless the length classifies as zero               (Ray.inl:74-85)
    const float len2 = __fsqrt_rn(dot_rn(r.ax, r.ay, r.az, r.ax, r.ay, r.az));
    if (len2 != 0.0f) { r.bx = __fdiv_rn(r.ax, len2); r.by = __fdiv_rn(r.ay, len2); r.bz = __fdiv_rn(r.az, len2); }
    else { r.bx = 0.0f; r.by = 0.0f; r.bz = 0.0f; }
    return r;
}

// Ray::intersect (src/Ray.cxx:72-125) on pre-packed per-triangle constants (prepack.h: ExactTri).
// A = {e2.xyz, e1.x}  B = {e1.yz, tvec.xy}  C = {tvec.z, qvec.xyz}
__device__ __forceinline__ bool exact_intersect(const RayDirs& r, const float4 A, const float4 B, const float4 C,
                                                const float e2q, float& t)
{
    const float pvx = __fsub_rn(__fmul_rn(r.by, A.z), __fmul_rn(r.bz, A.y));   // pvec = dir x edge2   (:90)
    const float pvy = __fsub_rn(__fmul_rn(r.bz, A.x), __fmul_rn(r.bx, A.z));
    const float pvz = __fsub_rn(__fmul_rn(r.bx, A.y), __fmul_rn(r.by, A.x));
    const float det = dot_rn(A.w, B.x, B.y, pvx, pvy, pvz);                      // edge1 . pvec        (:93)
    if (det == 0.0f) return false;                                              // FP_ZERO only        (:94)
    const float inv = __fdiv_rn(1.0f, det);                                     // float(1.0/det)      (:99)
    const float u = __fmul_rn(dot_rn(B.z, B.w, C.x, pvx, pvy, pvz), inv);       // (:105)
    if (u < 0.0f || u > 1.0f) return false;                                     // (:106)
    const float v = __fmul_rn(dot_rn(r.bx, r.by, r.bz, C.y, C.z, C.w), inv);    // dir . qvec          (:115)
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;                       // (:116)
    t = __fmul_rn(e2q, inv);                                                    // (edge2 . qvec)*inv  (:122)
    return true;
}

// lbuffer.cxx:788-793 for one accepted pair
__device__ __forceinline__ void accumulate_hit(const RayDirs& r, const float4 Dn, float t,
                                               float& dist, int& ssum, uint32_t& hits)
{
    if ((double)t > 0.0000001) {
        const float dn = dot_rn(r.ax, r.ay, r.az, Dn.y, Dn.z, Dn.w);            // direction . normal
        const int sign = (0.0f < dn) - (dn < 0.0f);                             // signum            (:729)
        dist = __fadd_rn(dist, __fmul_rn((float)sign, t));                      // distance += sign*t
        ssum += sign;
        ++hits;
    }
}

// lbuffer.cxx:798-809
__device__ __forceinline__ float attenuate(float L, float mu, float dist, int ssum, double length_to_cm)
{
    if (L == -1.0f) return L;                                                   // (:772) later meshes skipped
    if (ssum != 0) return -1.0f;
    return (float)((double)L * exp(-((double)mu * ((double)dist * length_to_cm))));
}

// ---- mbarrier / bulk-copy (TMA 1-D) plumbing ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init()
{ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}

// CTA -> work mapping.  CTAs are dispatched in ascending linear block index and the kernel ends when the slowest of the
// LAST ones does: CTAs over the object run the slow path and take up to twice as long as CTAs over empty space, which
// costs a fixed ~0.2 ms tail per launch (15 % of a rank's share at 8 GPUs).  The reference's camera centres the
// scene's bounding box in the view, so the remap below walks block rows and tile columns from the centre of the grid
// outwards -- the last CTAs are the cheap ones at the border -- and keeps the z-slices of one tile adjacent.
struct CtaCoord { uint32_t x, y, z; };
__device__ __forceinline__ uint32_t centre_out(uint32_t i, uint32_t n)
{
    const uint32_t mid = n >> 1, k = (i + 1) >> 1;
    return (i & 1u) ? mid - k : mid + k;
}
__device__ __forceinline__ CtaCoord cta_coord()
{
    const uint32_t gx = gridDim.x, gy = gridDim.y, gz = gridDim.z;
    const uint32_t L = blockIdx.x + gx * (blockIdx.y + gy * blockIdx.z);
    const uint32_t t = L / gz;
    CtaCoord c;
    c.z = L - t * gz;
    const uint32_t yi = t / gx, xi = t - yi * gx;
    c.x = centre_out(xi, gx); c.y = centre_out(yi, gy);
    return c;
}

// L output: the launch's own frame plus the extra (peer) frames of a fused multi-GPU assembly.  When L is an NVSwitch
// multicast mapping the store must be a multimem.st: the switch replicates it into every GPU's frame.  In halo mode
// (L_halo = repair depth) only the rows the neighbouring ranks' repair stencils reach are sent, and only to those two
// ranks: the first L_halo rows of a 16-row block to the owner of the block above, the last L_halo rows to the owner of
// the block below (with all frames replicated to all GPUs, the NVLink receive side bounds an 8-GPU frame).
__device__ __forceinline__ uint32_t halo_targets(const RenderArgs& a, uint32_t row)
{
    const uint32_t rel = row - a.row_begin, lrow = rel & 15u, blk = rel >> 4;
    const uint32_t last = (a.row_end - a.row_begin - 1) >> 4;
    return ((lrow < (uint32_t)a.L_halo && blk > 0) ? 1u : 0u) | ((lrow + (uint32_t)a.L_halo >= 16u && blk < last) ? 2u : 0u);
}
__device__ __forceinline__ void store_L(const RenderArgs& a, size_t p, uint32_t row, float v)
{
    if (a.L_multicast) asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(a.L + p), "f"(v) : "memory");
    else a.L[p] = v;
    if (a.L_halo) {
        const uint32_t t = halo_targets(a, row);
        if (t & 1u) a.L_extra[0][p] = v;
        if (t & 2u) a.L_extra[1][p] = v;
    } else
        for (uint32_t i = 0; i < a.n_L_extra; ++i) a.L_extra[i][p] = v;
}
__device__ __forceinline__ void store_L4(const RenderArgs& a, size_t p, uint32_t row, float4 v)      // p % 4 == 0, frames 16-byte aligned
{
    if (a.L_multicast)
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.L + p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else *reinterpret_cast<float4*>(a.L + p) = v;
    if (a.L_halo) {
        const uint32_t t = halo_targets(a, row);
        if (t & 1u) *reinterpret_cast<float4*>(a.L_extra[0] + p) = v;
        if (t & 2u) *reinterpret_cast<float4*>(a.L_extra[1] + p) = v;
    } else
        for (uint32_t i = 0; i < a.n_L_extra; ++i) *reinterpret_cast<float4*>(a.L_extra[i] + p) = v;
}

// Three-strip box test (filter proof item 7): may any pixel centre of the box hit the triangle?
constexpr float kBoxInflate = 1.0000038f;      // 1 + 2^-18

__device__ __forceinline__ bool box_may_hit(const float4 fa, const float4 fb, const float4 ga, const float4 gb,
                                            float xc, float yc, float hx, float hy)
{
    const float Dc = fabsf(fmaf(fa.w, xc, fmaf(fb.x, yc, fb.y)));
    const float rD = fmaf(fabsf(fa.w), hx, fabsf(fb.x) * hy);
    const float Y1 = fabsf(fmaf(fa.x, xc, fmaf(fa.y, yc, fa.z)));
    const float Y2 = fabsf(fmaf(ga.x, xc, fmaf(ga.y, yc, ga.z)));
    const float Y3 = fabsf(fmaf(gb.x, xc, fmaf(gb.y, yc, gb.z)));
    const float s1 = fmaf(fabsf(fa.x), hx, fmaf(fabsf(fa.y), hy, fmaf(2.0f, fb.z, rD))) * kBoxInflate;
    const float s2 = fmaf(fabsf(ga.x), hx, fmaf(fabsf(ga.y), hy, fmaf(2.0f, ga.w, rD))) * kBoxInflate;
    const float s3 = fmaf(fabsf(gb.x), hx, fmaf(fabsf(gb.y), hy, fmaf(2.0f, gb.w, rD))) * kBoxInflate;
    return !(Y1 - Dc > s1) && !(Y2 - Dc > s2) && !(Y3 - Dc > s3);          // NaN survives
}

// ------------------------------------------------------------------ exact kernel
// One pixel per thread, warp footprint 8 x 4 pixels; triangle constants are read with
// warp-uniform 128-bit loads (L1/L2 resident).  AUDIT additionally checks the filters.
template <bool AUDIT>
__global__ void __launch_bounds__(128) lbuffer_exact_kernel(const __grid_constant__ RenderArgs a)
{
    const uint32_t col = blockIdx.x * 8 + (threadIdx.x & 7);
    const uint32_t lrow = threadIdx.x >> 3;
    const uint32_t row = a.row_begin + (blockIdx.y * a.blk_stride + a.blk_phase) * 16 + lrow;
    const uint32_t orow = a.compact ? blockIdx.y * 16 + lrow : row;      // output row (slot-local when compact)
    const uint32_t W = a.cam.W;
    if (col >= W || row >= a.row_end) return;
    const RayDirs r = make_ray(a.cam, row, col);
    const float x = (float)col + a.cam.xc, y = (float)row + a.cam.yc;
    float L = a.incident;
    unsigned long long n_viol = 0;
    float min_strip = INFINITY, min_box = INFINITY;                       // audit: smallest head-room seen on a hit
    float zbuf = INFINITY; int zid = -1;                                  // mode 1
    for (int m = 0; m < a.n_meshes; ++m) {
        const DevMesh M = a.meshes[m];
        float dist = 0.0f; int ssum = 0; uint32_t hits = 0;
        if (M.contributes) {
#pragma unroll 1
            for (uint32_t j = 0; j < M.n; ++j) {
                const float4 A = __ldg(M.exact + 4 * (size_t)j);
                const float4 B = __ldg(M.exact + 4 * (size_t)j + 1);
                const float4 C = __ldg(M.exact + 4 * (size_t)j + 2);
                const float4 Dn = __ldg(M.exact + 4 * (size_t)j + 3);
                float t;
                const bool hit = exact_intersect(r, A, B, C, Dn.x, t);
                if (hit) {
                    if (a.mode == 0) accumulate_hit(r, Dn, t, dist, ssum, hits);
                    else if (zbuf > t) { zbuf = t; zid = (int)(M.first + j); }     // main-omp.cxx:656-662
                }
                if (AUDIT) {
                    const float4 fa = __ldg(M.filt0 + 3 * (size_t)j), fb = __ldg(M.filt0 + 3 * (size_t)j + 1);
                    const float4 ga = __ldg(M.filt1 + 2 * (size_t)j), gb = __ldg(M.filt1 + 2 * (size_t)j + 1);
                    const float X1 = fmaf(fa.x, x, fmaf(fa.y, y, fa.z));
                    const float Dd = fmaf(fa.w, x, fmaf(fb.x, y, fb.y));
                    const float X2 = fmaf(ga.x, x, fmaf(ga.y, y, ga.z));
                    const float X3 = fmaf(gb.x, x, fmaf(gb.y, y, gb.z));
                    // the filtered kernel's box of this pixel (aligned box_px x box_py box of the launch, clamped to the
                    // frame): level-0 gate at the box centre against the packed R, and the three-strip box test
                    const uint32_t c0 = col - col % a.box_px, r0 = row - lrow % a.box_py;
                    const uint32_t c1 = min(W, c0 + a.box_px) - 1, r1 = min(a.row_end, r0 + a.box_py) - 1;
                    const float bxc = 0.5f * (float)(c0 + c1) + a.cam.xc, byc = 0.5f * (float)(r0 + r1) + a.cam.yc;
                    const float Yc = fmaf(fa.x, bxc, fmaf(fa.y, byc, fa.z)), Dc = fmaf(fa.w, bxc, fmaf(fb.x, byc, fb.y));
                    const float4 fc = __ldg(M.filt0 + 3 * (size_t)j + 2);
                    const float Yc2 = fmaf(fc.x, bxc, fmaf(fc.y, byc, fc.z));
                    const bool gate = (fabsf(Yc) - fabsf(Dc) > fb.w) || (fabsf(Yc2) - fabsf(Dc) > fc.w) ||
                                      !box_may_hit(fa, fb, ga, gb, bxc, byc, 0.5f * (float)(c1 - c0), 0.5f * (float)(r1 - r0));
                    const bool reject = gate || (fabsf(X1) > fabsf(Dd) + fb.z) ||                                 // level 0
                                        (fabsf(X2) > fabsf(Dd) + ga.w) || (fabsf(X3) > fabsf(Dd) + gb.w);         // level 1
                    if (reject && hit) ++n_viol;
                    if (hit && fb.z > 0.0f) {
                        // head-room of the bounds on this HIT, in units of the rounding-error margin m: (|D| + m - |Y|) / m per
                        // strip (1 = the margin is untouched) and (R - (|Y(c)| - |D(c)|)) / 2m per level-0 box test (the box radii
                        // in R are exact geometry, 2m is its error budget; a hit on the strip's edge at a box corner gives 0.5)
                        const float aD = fabsf(Dd);
                        const float sp = fminf(fminf((aD + fb.z - fabsf(X1)) / fb.z, (aD + ga.w - fabsf(X2)) / ga.w), (aD + gb.w - fabsf(X3)) / gb.w);
                        const float sb = fminf((fb.w - (fabsf(Yc) - fabsf(Dc))) / (2.0f * fb.z), (fc.w - (fabsf(Yc2) - fabsf(Dc))) / (2.0f * ga.w));
                        min_strip = fminf(min_strip, sp); min_box = fminf(min_box, sb);
                    }
                }
            }
        }
        if (a.mode != 0) continue;
        L = attenuate(L, M.mu, dist, ssum, a.length_to_cm);
        if (m == 0) {
            const size_t p = (size_t)orow * W + col;
            if (a.dist) a.dist[p] = dist;
            if (a.ssum) a.ssum[p] = ssum;
            if (a.hits) a.hits[p] = hits;
        }
    }
    if (a.mode == 0) store_L(a, (size_t)(a.L_image ? row : orow) * W + col, row, L);
    else { a.dist[(size_t)orow * W + col] = zbuf; a.ssum[(size_t)orow * W + col] = zid; }
    if (AUDIT && a.counters) {
        if (n_viol) atomicAdd(a.counters + 1, n_viol);
        // non-negative floats order like their bit patterns; negative head-room is a violation and counted above
        if (min_strip < INFINITY) atomicMin(reinterpret_cast<unsigned int*>(a.counters + 2), __float_as_uint(fmaxf(min_strip, 0.0f)));
        if (min_box < INFINITY) atomicMin(reinterpret_cast<unsigned int*>(a.counters + 2) + 1, __float_as_uint(fmaxf(min_box, 0.0f)));
    }
}

// --------------------------------------------------------------- filtered kernel
// Brute force: every triangle meets every box of pixels, and every (pixel, triangle) pair is decided either by a
// proven-conservative reject or by the reference's exact sequence.  Four kernels:
//
// 1. lbuffer_filtered_kernel (the scan: > 90 % of the time).  Triangle-stationary: a thread holds the level-0 records of
//    K triangles; a CTA (128 threads, 128 K triangles: "chunk" z of the mesh, a strided sample -- triangle
//    (k 128 + t) n_chunks + z) walks the boxes of PX x PY pixels of its image region: the 16 / PY box rows of one 16-row
//    block (the block-cyclic unit of the multi-GPU split) x a range of box columns.  Box centres are warp-uniform
//    scalars computed from the loop counters, so the hot loop touches NO memory at all.
//    (The first version of this round streamed triangle tiles through shared memory with each thread owning boxes: a
//    broadcast LDS.128 still moves 512 B into the register file, 4 LSU cycles per warp, and the 3 loads per
//    (warp, triangle) bounded that kernel at 12 cycles per warp-triangle -- measured 13.7 -- with the issue slots half
//    idle: profiles/r02_sweep_tile_streaming_*.txt.)
//    Per (box, triangle): Y, Y2 (two of the triangle's three strips) and D are affine in the pixel coordinates; they are
//    evaluated at the box CENTRE and the box is rejected when |Y(c)| - |D(c)| > R or |Y2(c)| - |D(c)| > R2, R being the
//    record's pre-packed bound r_Y + r_D + 2 m for this box size (filter proof item 7).  One strip is an
//    infinite band across the image; two strips intersect in a parallelogram around the triangle, so survivors are rare
//    (3e-4 of the box tests at 2048^2).  3 FFMA + 2 FADD + 2 FSETP per test, one vote per two boxes.
//    Survivors {box, triangle} are staged in a per-warp shared-memory buffer and appended, 128 at a time, to a global
//    candidate list (one atomic per flush).  Nothing latency-bound runs in this kernel: when the exact tests ran here
//    too (a warp draining its own queue), the dependent L2 round trips of ~100 000 drain passes held warp slots for
//    0.5 ms of a 1.4 ms frame and, worse, 0.2 ms of every 0.33 ms share of an 8-GPU frame
//    (profiles/r02_sweep_scan_inline_drains.txt).
// 2. candidates_kernel, one list entry per thread: the full three-strip box test (the triangle's conservative
//    footprint), then the per-pixel three-strip tests over the pixels of the box; the surviving {pixel, triangle} pairs
//    (1.005 x the true hits) go to a second global list.
// 3. exact_tests_kernel, one pair per thread: ray generation with both normalisations, Ray::intersect's exact
//    sequence, sign test; accepted intersections are appended to the pixel's hit list through an atomic slot counter.
//    (2 and 3 are latency-bound chains of dependent L2 accesses: one item per thread and tens of thousands of threads
//    in flight hide what a per-warp queue inside the scan kernel could not.)
// 4. resolve_hits_kernel: sorts each pixel's list by triangle index and replays the reference's ordered float sum --
//    so the result does not depend on which thread or chunk found a hit: bit-identical to the serial loop.
// If a candidate list overflows (adversarial scenes; the host grows it for the next frame), resolve recomputes every
// pixel with the exact sequence: slow, never wrong.
// occupancy targets of the latency-bound exact-test and resolve kernels (CTAs per SM; the register cap follows: 40 registers,
// a few spills, 3 / 2 us faster; the candidates kernel lost 2 us the same way and keeps its 56 registers)
#ifndef SRT_RESOLVE_MIN_CTAS
#define SRT_RESOLVE_MIN_CTAS 6
#endif
#ifndef SRT_EXACT_MIN_CTAS
#define SRT_EXACT_MIN_CTAS 12
#endif
constexpr uint32_t kStageCap = 256, kStageFlush = 128;
constexpr uint32_t kNoHit = 0xffffffffu;       // empty list head / end of a pixel's hit list

// candidate entry: x = box row j (4 bits) | 16-row block of the launch (14) | box column i (14), y = triangle (27) | mesh (5)
__device__ __forceinline__ uint2 cand_pack(uint32_t j, uint32_t block, uint32_t i, uint32_t tri, uint32_t mesh)
{ return make_uint2(j | (block << 4) | (i << 18), tri | (mesh << 27)); }

#ifndef SRT_MIN_CTAS
#define SRT_MIN_CTAS 4
#endif

template <int PX, int PY, int K>
struct ScanGeom {
    static_assert(16 % PY == 0 && PX <= 64, "box rows tile a 16-row block");
    // boxes of more than 64 pixels are refined into 16 x 4 sub-boxes before the per-pixel stage (refine_kernel)
    static constexpr bool kBig = PX * PY > 64;
    static_assert(!kBig || (PX % 16 == 0 && PY % 4 == 0), "big boxes are whole 16 x 4 sub-boxes");
    static constexpr uint32_t kBoxRows = 16 / PY;
    static constexpr size_t kSmem = (size_t)K * 3 * 128 * sizeof(float4) + 16 + 4 * kStageCap * sizeof(uint2);
};

template <int PX, int PY, int K>
__global__ void __launch_bounds__(128, SRT_MIN_CTAS) lbuffer_filtered_kernel(const __grid_constant__ RenderArgs a)
{
    using G = ScanGeom<PX, PY, K>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* recs = reinterpret_cast<float4*>(smem_raw);                        // [K][3][128]: the CTA's current chunk, SoA
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)K * 3 * 128 * sizeof(float4));
    uint2* stage = reinterpret_cast<uint2*>(bar + 2) + (threadIdx.x >> 5) * kStageCap;       // this warp's survivors

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t W = a.cam.W;
    const uint32_t n_cols = (W + PX - 1) / PX;                                               // box columns of the frame
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    // the list counters are double-buffered: this launch counts into cand_count, which the previous launch of the path
    // left at zero, and leaves cand_count_next at zero for the next one (no memset between frames)
    if (tid < 4 && a.cand_count_next && blockIdx.x == 0) a.cand_count_next[tid] = 0;
    __syncthreads();
    uint32_t phase = 0, n1 = 0;              // n1: survivors staged by this warp (warp-uniform)
    const float x_first = 0.5f * (float)(PX - 1) + a.cam.xc;                                 // centre of box column 0

    // Persistent CTAs, static partition.  The work of the launch is the sequence of UNITS
    //   (chunk, 16-row block, box row, range of scan_cols box columns),  chunk-major,
    // every unit the same number of box tests (up to the frame's ragged edge); CTA g of G takes the contiguous range
    // [g T / G, (g + 1) T / G) of the T units, so all CTAs finish together whatever the share of the frame and the number
    // of resident CTAs (no partial last wave), and a CTA loads a new chunk at most twice.  Boxes over the object stage
    // survivors and cost more than boxes over empty space, so the blocks of a chunk are visited in a strided order
    // (block (t s) mod n, s coprime to n, ~0.62 n): the dozen blocks of a CTA's range sample the whole frame and every
    // CTA gets the same mix (with contiguous bands the CTAs over the object's centre finished 20 % late).
    const uint32_t parts_x = a.scan_parts_x;
    const uint64_t per_chunk = (uint64_t)a.n_blocks * G::kBoxRows * parts_x;
    const uint64_t total = per_chunk * a.scan_chunks;
    const uint64_t u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
    uint32_t gz = (uint32_t)(u0 / per_chunk);                                                // chunk, counted over the contributing meshes
    uint32_t part, j, bt, blk;                                                              // bt: position in the block order, blk: the block
    {
        uint64_t r = u0 - (uint64_t)gz * per_chunk;
        part = (uint32_t)(r % parts_x); r /= parts_x;
        j = (uint32_t)(r % G::kBoxRows); bt = (uint32_t)(r / G::kBoxRows);
        blk = (uint32_t)(((uint64_t)bt * a.scan_blk_stride) % a.n_blocks);
    }
    bool new_chunk = true, new_row = true, row_ok = false;
    uint32_t m = 0, z = 0, n_chunks = 1;     // mesh / chunk within the mesh / chunks of that mesh

    // append the warp's staged survivors to the global candidate list: one atomic, coalesced 8-byte stores
    auto flush = [&]() {
        __syncwarp();
        uint32_t base = 0;
        uint2* const list = G::kBig ? a.big : a.cand;
        const uint32_t cap = G::kBig ? a.big_cap : a.cand_cap;
        if (lane == 0) base = atomicAdd(G::kBig ? a.big_count : a.cand_count, n1);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t e = lane; e < n1; e += 32)
            if (base + e < cap) list[base + e] = stage[e];                                  // overflow: resolve falls back to the exact sequence
        __syncwarp();
        n1 = 0;
    };

    // the thread's K triangles: x slopes and bounds, and the row parts of Y, Y2, D folded with the row's yc -- the inner
    // loop lives on these 8 values per triangle (re-read from shared memory once per box row: 12 LDS per ~100 boxes)
    float bx[K], bd[K], R[K], bx2[K], R2[K], by[K], dy[K], by2[K];
    // level-0 test of the thread's K triangles against the box with centre xc: may any of them touch it?
    auto scan = [&](float xc) {
        bool surv = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float D = fabsf(fmaf(bd[k], xc, dy[k]));
            const float Y = fmaf(bx[k], xc, by[k]), Y2 = fmaf(bx2[k], xc, by2[k]);
            const bool r1 = fabsf(Y) - D > R[k], r2 = fabsf(Y2) - D > R2[k];          // NaN survives
            surv |= !(r1 | r2);                                                      // bitwise: no short-circuit branches
        }
        return surv;
    };
    // stage the survivors of box (j, i) -- rare: one ballot per k.  The tests are re-evaluated on an opaque copy
    // of xc so that the compiler does not carry the K predicates of every box through the hot loop.
    auto push = [&](uint32_t i, float xc) {
        asm volatile("" : "+f"(xc));
        if (n1 + 32 * K > kStageCap) flush();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float D = fabsf(fmaf(bd[k], xc, dy[k]));
            const float Y = fmaf(bx[k], xc, by[k]), Y2 = fmaf(bx2[k], xc, by2[k]);
            const bool sv = !((fabsf(Y) - D > R[k]) | (fabsf(Y2) - D > R2[k]));
            const unsigned bal = __ballot_sync(0xffffffffu, sv);
            if (sv) stage[n1 + __popc(bal & ((1u << lane) - 1u))] =
                cand_pack(j, blk, i, ((uint32_t)k * 128 + tid) * n_chunks + z, m);
            n1 += __popc(bal);
        }
        if (n1 >= kStageFlush) flush();
    };

#pragma unroll 1
    for (uint64_t u = u0; u < u1; ++u) {
        if (new_chunk) {
            // chunk gz of the launch -> (mesh, chunk of the mesh): meshes that add no path length hold no chunks
            uint32_t rest = gz;
            const DevMesh* Mp = a.meshes;
            for (m = 0;; ++m, ++Mp) {
                const uint32_t nc = (Mp->contributes && Mp->n) ? Mp->n_chunks : 0u;
                if (rest < nc) { n_chunks = nc; break; }
                rest -= nc;
            }
            z = rest;
            // the chunk's level-0 records: one bulk copy (TMA) into shared memory (SoA in the chunk: consecutive lanes
            // read consecutive float4)
            __syncthreads();                                                                 // recs of the previous chunk are consumed
            if (tid == 0) {
                mbar_expect_tx(bar, (uint32_t)(K * 3 * 128 * sizeof(float4)));
                bulk_g2s(recs, Mp->scan + (size_t)z * K * 3 * 128, (uint32_t)(K * 3 * 128 * sizeof(float4)), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            new_chunk = false; new_row = true;
        }
        if (new_row) {
            const uint32_t br0 = a.row_begin + (blk * a.blk_stride + a.blk_phase) * 16 + j * PY;   // first pixel row of the box row
            row_ok = br0 < a.row_end;
            const uint32_t br1 = min(a.row_end, br0 + PY) - 1;                                // boxes are clamped to the frame
            const float yc = 0.5f * (float)(br0 + br1) + a.cam.yc;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4 fa = recs[(k * 3 + 0) * 128 + tid], fb = recs[(k * 3 + 1) * 128 + tid], fc = recs[(k * 3 + 2) * 128 + tid];
                bx[k] = fa.x; bd[k] = fa.w; R[k] = fb.w; bx2[k] = fc.x; R2[k] = fc.w;
                by[k] = fmaf(fa.y, yc, fa.z); dy[k] = fmaf(fb.x, yc, fb.y); by2[k] = fmaf(fc.y, yc, fc.z);
            }
            new_row = false;
        }
        if (row_ok) {
            const uint32_t c_begin = part * a.scan_cols, c_end = min(n_cols, c_begin + a.scan_cols);
            // full boxes: the centre advances by PX (exact in binary32: half-integers far below 2^23)
            const uint32_t c_full = min(c_end, W / PX);
            float xc = fmaf((float)c_begin, (float)PX, x_first);
            // two boxes per vote: 2 K independent chains per warp, half the vote / branch overhead
            uint32_t i = c_begin;
#pragma unroll 1
            for (; i + 1 < c_full; i += 2, xc += (float)(2 * PX)) {
                const float xd = xc + (float)PX;
                const bool s0 = scan(xc), s1 = scan(xd);
                if (__any_sync(0xffffffffu, s0 | s1)) {
                    if (__any_sync(0xffffffffu, s0)) push(i, xc);
                    if (__any_sync(0xffffffffu, s1)) push(i + 1, xd);
                }
            }
            if (i < c_full && __any_sync(0xffffffffu, scan(xc))) push(i, xc);
            if (c_full < c_end) {                               // the last box of the row is clamped to the frame
                const float xl = 0.5f * (float)(c_full * PX + W - 1) + a.cam.xc;
                if (__any_sync(0xffffffffu, scan(xl))) push(c_full, xl);
            }
        }
        // next unit: column range, then box row, then block, then chunk
        if (++part == parts_x) {
            part = 0; new_row = true;
            if (++j == G::kBoxRows) {
                j = 0;
                blk += a.scan_blk_stride; if (blk >= a.n_blocks) blk -= a.n_blocks;
                if (++bt == a.n_blocks) { bt = 0; blk = 0; ++gz; new_chunk = true; }
            }
        }
    }
    if (n1) flush();
}

// ------------------------------------------------------------------ refine kernel (big-box shapes)
// One {big box, triangle} survivor of the scan per thread (grid-stride): the three-strip box test (the triangle's
// conservative footprint, DESIGN.md filter proof item 7) on each 16 x 4 sub-box of the big box; surviving
// {sub-box, triangle} entries are appended to the box list of candidates_kernel<16, 4> (one atomic per warp and
// sub-box position).  The scan's two-strip test on a big box lets ~10 x more area through than the triangle covers;
// this stage cuts it back before the per-pixel tests.
template <int PX, int PY>
__global__ void __launch_bounds__(128) refine_kernel(const __grid_constant__ RenderArgs a)
{
    constexpr uint32_t SX = PX / 16, SY = PY / 4;                              // sub-boxes per big box
    static_assert(SX * SY <= 32, "one bit per sub-box");
    const uint32_t lane = threadIdx.x & 31u, W = a.cam.W;
    const uint32_t n = min(*a.big_count, a.big_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {      // warp-uniform trip count
        const uint32_t e = base + lane;
        const bool live = e < n;
        uint32_t tm = 0, blk = 0, jb = 0, ib = 0, row0 = 0;
        float4 fa = make_float4(0, 0, 0, 0), fb = fa, ga = fa, gb = fa;
        if (live) {
            const uint2 ent = a.big[e];
            tm = ent.y;
            const DevMesh* Mp = a.meshes + (tm >> 27);
            const uint32_t tri = tm & 0x7ffffffu;
            jb = ent.x & 15u; blk = (ent.x >> 4) & 0x3fffu; ib = ent.x >> 18;
            row0 = a.row_begin + (blk * a.blk_stride + a.blk_phase) * 16 + jb * PY;
            const float4* f0 = Mp->filt0 + 3 * (size_t)tri; const float4* f1 = Mp->filt1 + 2 * (size_t)tri;
            fa = __ldg(f0); fb = __ldg(f0 + 1); ga = __ldg(f1); gb = __ldg(f1 + 1);
        }
        // which sub-boxes may the triangle touch?  (no atomics in this loop: one per warp below)
        uint32_t mask = 0;
        if (live) {
#pragma unroll
            for (uint32_t sy = 0; sy < SY; ++sy) {
#pragma unroll
                for (uint32_t sx = 0; sx < SX; ++sx) {
                    const uint32_t r0 = row0 + sy * 4, c0 = (ib * SX + sx) * 16;
                    if (r0 < a.row_end && c0 < W) {
                        const uint32_t c1 = min(W, c0 + 16) - 1, r1 = min(a.row_end, r0 + 4) - 1;     // clamped to the frame
                        if (box_may_hit(fa, fb, ga, gb, 0.5f * (float)(c0 + c1) + a.cam.xc, 0.5f * (float)(r0 + r1) + a.cam.yc,
                                        0.5f * (float)(c1 - c0), 0.5f * (float)(r1 - r0)))
                            mask |= 1u << (sy * SX + sx);
                    }
                }
            }
        }
        // append the warp's sub-boxes: one atomic, each lane writes its own
        const uint32_t mine = (uint32_t)__popc(mask);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        uint32_t at = 0;
        if (lane == 31) at = atomicAdd(a.cand_count, total);
        at = __shfl_sync(0xffffffffu, at, 31) + incl - mine;
        while (mask) {
            const uint32_t k = (uint32_t)__ffs((int)mask) - 1;
            mask &= mask - 1;
            if (at < a.cand_cap) a.cand[at] = cand_pack(jb * SY + k / SX, blk, ib * SX + k % SX, tm & 0x7ffffffu, tm >> 27);
            ++at;
        }
    }
}

// ------------------------------------------------------------------ candidates kernel
// image row -> output row (slot-local when the launch is compact)
__device__ __forceinline__ uint32_t out_row(const RenderArgs& a, uint32_t row)
{
    if (!a.compact) return row;
    const uint32_t rel = row - a.row_begin;
    return ((rel >> 4) - a.blk_phase) / a.blk_stride * 16 + (rel & 15u);
}

// Stage 2 of the filtered path, one list entry {box, triangle} per thread (grid-stride): records from L2, the
// three-strip box test (the triangle's conservative footprint), then the per-pixel three-strip tests over the pixels
// of the box with the records in registers.  The surviving {pixel, triangle} pairs -- the candidates of the exact test,
// 1.005 x the true hits -- are appended to a second global list (one atomic per warp).
template <int PX, int PY>
__global__ void __launch_bounds__(128) candidates_kernel(const __grid_constant__ RenderArgs a)
{
    static_assert(PX * PY <= 64, "the per-entry pixel mask is 64 bits");
    const uint32_t lane = threadIdx.x & 31u, W = a.cam.W;
    const uint32_t n = min(a.cand_count[0], a.cand_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {      // warp-uniform trip count
        const uint32_t e = base + lane;
        unsigned long long mask = 0;
        uint32_t tm = 0, row0 = 0, col0 = 0;
        if (e < n) {
            const uint2 ent = a.cand[e];
            const uint32_t box_x = ent.x;
            tm = ent.y;
            const uint32_t tri = tm & 0x7ffffffu;
            const DevMesh* Mp = a.meshes + (tm >> 27);
            row0 = a.row_begin + ((((box_x >> 4) & 0x3fffu) * a.blk_stride + a.blk_phase) * 16) + (box_x & 15u) * PY;
            col0 = (box_x >> 18) * PX;
            const uint32_t c1 = min(W, col0 + PX) - 1, r1 = min(a.row_end, row0 + PY) - 1;      // listed boxes hold pixels
            const float4* f0 = Mp->filt0 + 3 * (size_t)tri; const float4* f1 = Mp->filt1 + 2 * (size_t)tri;
            const float4 fa = __ldg(f0), fb = __ldg(f0 + 1), ga = __ldg(f1), gb = __ldg(f1 + 1);
            if (box_may_hit(fa, fb, ga, gb, 0.5f * (float)(col0 + c1) + a.cam.xc, 0.5f * (float)(row0 + r1) + a.cam.yc,
                            0.5f * (float)(c1 - col0), 0.5f * (float)(r1 - row0))) {
                const float x0 = (float)col0 + a.cam.xc;
#pragma unroll
                for (int ky = 0; ky < PY; ++ky) {
                    const float y = (float)(row0 + ky) + a.cam.yc;
                    const float bd = fmaf(fb.x, y, fb.y), b1 = fmaf(fa.y, y, fa.z), b2 = fmaf(ga.y, y, ga.z), b3 = fmaf(gb.y, y, gb.z);
                    unsigned rowmask = 0;
#pragma unroll
                    for (int kx = 0; kx < PX; ++kx) {
                        const float x = x0 + (float)kx;                                 // exact: (half-)integers far below 2^23
                        const float Dd = fabsf(fmaf(fa.w, x, bd));
                        const float Y1 = fmaf(fa.x, x, b1), Y2 = fmaf(ga.x, x, b2), Y3 = fmaf(gb.x, x, b3);
                        // reject iff |Y_i| > fl(Dd + m_i) for some strip  <=>  fl(fl(Dd + m_i) - |Y_i|) < 0 (the sign of a rounded
                        // difference is the sign of the exact one); fminf drops NaN operands, and an all-NaN pixel survives as
                        // with the comparisons.  Written this way the work sits on the FMA pipe (6 FADD) instead of the ALU
                        // pipe the kernel was bound by (3 FSETP + 3 SEL per pixel).
                        const float t = fminf(fminf((Dd + fb.z) - fabsf(Y1), (Dd + ga.w) - fabsf(Y2)), (Dd + gb.w) - fabsf(Y3));
                        if (!(t < 0.0f)) rowmask |= 1u << kx;
                    }
                    if (col0 + PX > W) rowmask &= (1u << (W - col0)) - 1u;              // pixels beyond the last column
                    if (row0 + ky < a.row_end) mask |= (unsigned long long)rowmask << (ky * PX);
                }
            }
        }
        // append the warp's candidates: one atomic, each lane writes its own pixels
        const uint32_t mine = (uint32_t)__popcll(mask);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        uint32_t at = 0;
        if (lane == 31) at = atomicAdd(a.cand_count + 1, total);
        at = __shfl_sync(0xffffffffu, at, 31) + incl - mine;
        // (32-bit halves: the 64-bit find-first-set / clear-lowest pair costs three times the instructions, and this loop runs
        // for as long as the lane with the most candidates)
        const uint32_t pix0 = row0 * W + col0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t m = half ? (uint32_t)(mask >> 32) : (uint32_t)mask;
            while (m) {
                const uint32_t k = (uint32_t)__ffs((int)m) - 1 + 32u * half;
                m &= m - 1;
                if (at < a.pix_cap) a.pix[at] = make_uint2(pix0 + (k / PX) * W + k % PX, tm);   // overflow: resolve falls back
                ++at;
            }
        }
    }
}

// Stage 3, one candidate {pixel, triangle} per thread (grid-stride): ray generation with both normalisations,
// Ray::intersect's exact sequence, sign test; accepted intersections are appended to the pixel's hit list through an
// atomic slot counter.
__global__ void __launch_bounds__(128, SRT_EXACT_MIN_CTAS) exact_tests_kernel(const __grid_constant__ RenderArgs a)
{
    const uint32_t W = a.cam.W;
    const uint32_t n = min(a.cand_count[1], a.pix_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    if (a.counters && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.counters, (unsigned long long)n);   // pairs that run the exact sequence
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const uint2 ent = a.pix[e];
        const uint32_t tri = ent.y & 0x7ffffffu;
        const DevMesh* Mp = a.meshes + (ent.y >> 27);
        const float4* exact = Mp->exact + 4 * (size_t)tri;
        const float4 A = __ldg(exact), B = __ldg(exact + 1), C = __ldg(exact + 2), Dn = __ldg(exact + 3);
        const uint32_t row = ent.x / W, col = ent.x - row * W;
        const RayDirs r = make_ray(a.cam, row, col);
        float tt;
        if (!exact_intersect(r, A, B, C, Dn.x, tt)) continue;
        float val = tt;                                                   // mode 1: z-buffer candidate (any sign of t)
        if (a.mode == 0) {
            if (!((double)tt > 0.0000001)) continue;                       // lbuffer.cxx:788
            const float dn = dot_rn(r.ax, r.ay, r.az, Dn.y, Dn.z, Dn.w);   // direction . normal, pre-Ray direction
            val = __fmul_rn((float)((0.0f < dn) - (dn < 0.0f)), tt);       // sign * t
        } else if (a.mode == 2 && !((double)tt > 0.0000001)) continue;     // sort mode keeps t itself (main.cxx:687)
        // push onto the pixel's list: the candidate's own slot e becomes the list node {triangle, value} (the pixel index
        // it held is no longer needed), linked through hit_next; one atomic exchange on the pixel's head
        const size_t p = (size_t)out_row(a, row) * W + col;
        a.pix[e] = make_uint2(Mp->first + tri, __float_as_uint(val));
        a.hit_next[e] = atomicExch(a.hit_head + p, e);
    }
}

// ------------------------------------------------------------------ hit-list resolve
// Second pass of the filtered kernel: one thread per output pixel sorts the pixel's hit list by scene-wide triangle
// index (insertion sort, lists are a handful of entries) and replays the reference's per-mesh accumulation and
// attenuation (lbuffer.cxx:768-809), or the z-buffer update of the shaded drivers (main-omp.cxx:656-662), in ascending
// triangle order.  It also resets the pixel's slot counter for the next render.  A list that overflowed its capacity is
// recomputed from scratch with the exact sequence over all triangles (rare; keeps the result exact).
// std::exp(float) of the reference's build == glibc's expf (sysdeps/ieee754/flt-32/e_expf.c; the exp2f algorithm of the
// ARM optimized routines): x N / ln2 = k + r in double, 2^(k/N) from a 32-entry table, a cubic in r, one narrowing.  Its
// result is the correctly rounded float in all but ~0.06 % of the arguments; restating the algorithm (table entry i =
// bits(2^(i/32)) - (i << 47), coefficients as published) reproduces those too -- checked against the libm of the build
// container on 200 000 arguments (tools/check_expf.py).
__device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

__device__ __forceinline__ float expf_glibc(float x)
{
    const uint32_t abstop = (__float_as_uint(x) >> 20) & 0x7ffu;
    if (abstop >= 0x42bu) {                                   // |x| >= 88 or NaN
        if (__float_as_uint(x) == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8u) return x + x;
        if (x > 88.72283172607421875f) return INFINITY;       // 0x1.62e42ep6f
        if (x < -103.972076416015625f) return 0.0f;           // -0x1.9fe368p6f
    }
    const double N = 32.0;
    const double z = __dmul_rn(0x1.71547652b82fep+0 * N, (double)x);
    double kd = __dadd_rn(z, 0x1.8p+52);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, 0x1.8p+52);
    const double r = __dsub_rn(z, kd);
    const unsigned long long t = kExp2fTab[ki % 32] + (ki << 47);
    const double s = __longlong_as_double((long long)t);
    const double c0 = 0x1.c6af84b912394p-5 / N / N / N, c1 = 0x1.ebfce50fac4f3p-3 / N / N, c2 = 0x1.62e42ff0c52d6p-1 / N;
    const double p = __dadd_rn(__dmul_rn(c0, r), c1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(c2, r), 1.0);
    y = __dadd_rn(__dmul_rn(p, r2), y);
    y = __dmul_rn(y, s);
    return (float)y;
}

constexpr uint32_t kLocalHits = 32;            // hits of one pixel the resolve pass sorts in a local array
constexpr uint32_t kSortOverflowCap = 256;     // sort mode: intersections of one ray the overflow path can order

// Sort-based X-ray value of one pixel (src/main.cxx:700-741): the ray's intersection distances t (mesh 0, t > 1e-7) are
// sorted, exit - entry pairs are summed in float, an odd count counts as no path (:714-718), then
// `distance = distance * 0.1` (double product, narrowed) and `80.000f * exp(-(0.3971f * distance))` with std::exp(float)
// (expf_glibc above).
__device__ float sorted_xray_value(float* t, uint32_t n, float incident, float mu, double length_to_cm, float& path)
{
    float distance = 0.0f;
    if (n > 0 && (n & 1u) == 0) {
        for (uint32_t i = 1; i < n; ++i) {                  // std::sort: ascending; ties are interchangeable values
            const float key = t[i];
            uint32_t j = i;
            while (j > 0 && t[j - 1] > key) { t[j] = t[j - 1]; --j; }
            t[j] = key;
        }
        for (uint32_t i = 0; i < n; i += 2) distance = __fadd_rn(distance, __fsub_rn(t[i + 1], t[i]));    // :706-708
    }
    path = distance;
    distance = (float)((double)distance * length_to_cm);                                              // :725
    return __fmul_rn(incident, expf_glibc(-__fmul_rn(mu, distance)));                                  // :739
}

// One pixel's outputs from its hit list (head = first node, kNoHit = empty): L (modes 0, 2), path length / z value,
// sign sum / triangle index, hit count.
struct PixelOut { float L, dist; int ssum; uint32_t hits; };

__device__ __forceinline__ PixelOut resolve_list(const RenderArgs& a, uint32_t head)
{
    float L = a.incident;
    float dist0 = 0.0f; int ssum0 = 0; uint32_t hits0 = 0;   // mode 1: z-buffer value / triangle index
    if (a.mode == 1) {
        // closest hit: strict `z > t` in ascending triangle order == smallest t, lowest triangle index among equals
        dist0 = INFINITY; ssum0 = -1;
        for (uint32_t i = head; i != kNoHit; i = a.hit_next[i]) {
            const uint2 e = a.pix[i];
            const float t = __uint_as_float(e.y);
            if (dist0 > t || (dist0 == t && (int)e.x < ssum0)) { dist0 = t; ssum0 = (int)e.x; }
        }
        return { L, dist0, ssum0, hits0 };
    }
    // The pixel's hits in the reference's order: ascending triangle index (mode 0: the float sum is order
    // dependent) resp. ascending t (mode 2).
    // Fast path, lists of up to 8 hits (nearly all): the nodes are sorted in REGISTERS by an unrolled insertion
    // network -- no indexed local array, which costs a local-memory round trip per access.
    {
        uint2 r[8];
        uint32_t n8 = 0, i = head;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (i != kNoHit) {
                r[s] = a.pix[i]; i = a.hit_next[i]; n8 = s + 1;
#pragma unroll
                for (int j = s; j > 0; --j) {
                    const bool sw = (a.mode == 2 ? r[j - 1].y : r[j - 1].x) > (a.mode == 2 ? r[j].y : r[j].x);
                    const uint2 lo = sw ? r[j] : r[j - 1], hi = sw ? r[j - 1] : r[j];
                    r[j - 1] = lo; r[j] = hi;
                }
            }
        }
        if (i == kNoHit && n8 > 0) {
            if (a.mode == 2) {
                // main.cxx:700-741 (sorted_xray_value): pairs (t[i], t[i + 1]) in ascending order
                float distance = 0.0f;
                if ((n8 & 1u) == 0) {
#pragma unroll
                    for (int s = 0; s < 8; s += 2)
                        if ((uint32_t)s < n8) distance = __fadd_rn(distance, __fsub_rn(__uint_as_float(r[s + 1].y), __uint_as_float(r[s].y)));
                }
                dist0 = distance;
                distance = (float)((double)distance * a.length_to_cm);
                L = __fmul_rn(a.incident, expf_glibc(-__fmul_rn(a.meshes[0].mu, distance)));
                return { L, dist0, 0, n8 };
            }
            for (int m = 0; m < a.n_meshes; ++m) {
                const DevMesh M = a.meshes[m];
                float dist = 0.0f; int ssum = 0; uint32_t hits = 0;
                if (M.contributes) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {                    // ascending triangle order; mesh ranges ascend with m
                        if ((uint32_t)s < n8 && r[s].x >= M.first && r[s].x < M.first + M.n) {
                            const float st = __uint_as_float(r[s].y);
                            dist = __fadd_rn(dist, st);              // distance += sign * t
                            ssum += (0.0f < st) - (st < 0.0f);
                            ++hits;
                        }
                    }
                }
                L = attenuate(L, M.mu, dist, ssum, a.length_to_cm);
                if (m == 0) { dist0 = dist; ssum0 = ssum; hits0 = hits; }
            }
            return { L, dist0, ssum0, hits0 };
        }
    }
    // Longer lists: up to kLocalHits entries are insertion-sorted into a local array while the chain is walked; beyond
    // that they are replayed window by window (below).
    uint2 e[kLocalHits];
    uint32_t n = 0;
    for (uint32_t i = head; i != kNoHit; i = a.hit_next[i], ++n) {
        if (n >= kLocalHits) continue;
        const uint2 key = a.pix[i];
        const uint32_t kk = a.mode == 2 ? key.y : key.x;            // positive floats order like their bits
        uint32_t j = n;
        while (j > 0 && (a.mode == 2 ? e[j - 1].y : e[j - 1].x) > kk) { e[j] = e[j - 1]; --j; }
        e[j] = key;
    }
    // visit entry number `i` (0-based; the callers ask for 0, 1, 2, ... in order): from the array when the whole list
    // fits, else from a WINDOW of the next kLocalHits entries in order, refilled by one walk of the chain that keeps the
    // kLocalHits smallest (key, node) pairs above the last one visited -- n / kLocalHits walks per list instead of the n
    // a plain selection takes (the poles of the 998 000-triangle sphere put ~1 400 hits on some rays).
    unsigned long long last = 0;                                    // (key, node) + 1 of the last entry of the previous window
    unsigned long long wk[kLocalHits];                              // the window's (key, node) + 1, ascending; values in e[]
    uint32_t win_base = 0, win_n = 0;
    auto entry = [&](uint32_t i) -> uint2 {
        if (n <= kLocalHits) return e[i];
        if (i >= win_base + win_n) {
            win_base += win_n; win_n = 0;
            for (uint32_t c = head; c != kNoHit; c = a.hit_next[c]) {
                const uint2 v = a.pix[c];
                const unsigned long long k = (((unsigned long long)(a.mode == 2 ? v.y : v.x)) << 32 | c) + 1;
                if (k <= last || (win_n == kLocalHits && k >= wk[kLocalHits - 1])) continue;
                uint32_t j = win_n < kLocalHits ? win_n++ : kLocalHits - 1;      // full: the largest drops out
                while (j > 0 && wk[j - 1] > k) { wk[j] = wk[j - 1]; e[j] = e[j - 1]; --j; }
                wk[j] = k; e[j] = v;
            }
            last = wk[win_n - 1];
        }
        return e[i - win_base];
    };
    if (a.mode == 2) {
        // main.cxx:700-741 (sorted_xray_value), streamed: pairs (t[i], t[i + 1]) in ascending order
        float distance = 0.0f;
        if (n > 0 && (n & 1u) == 0)
            for (uint32_t i = 0; i < n; i += 2) {
                const float t0 = __uint_as_float(entry(i).y), t1 = __uint_as_float(entry(i + 1).y);
                distance = __fadd_rn(distance, __fsub_rn(t1, t0));
            }
        dist0 = distance;
        distance = (float)((double)distance * a.length_to_cm);
        L = __fmul_rn(a.incident, expf_glibc(-__fmul_rn(a.meshes[0].mu, distance)));
        ssum0 = 0; hits0 = n;
    } else if (n == 0) {
        // no hit in any mesh: L *= exp(-(mu * (0 * 0.1))) == L exactly, unless mu or the unit factor is not finite
        for (int m = 0; m < a.n_meshes; ++m) {
            const float mu = a.meshes[m].mu;
            if (!(isfinite(mu) && isfinite(a.length_to_cm))) L = attenuate(L, mu, 0.0f, 0, a.length_to_cm);
        }
    } else {
        uint32_t i = 0;
        uint2 cur = entry(0);
        for (int m = 0; m < a.n_meshes; ++m) {
            const DevMesh M = a.meshes[m];
            float dist = 0.0f; int ssum = 0; uint32_t hits = 0;
            if (M.contributes) {
                while (i < n && cur.x < M.first + M.n) {
                    const float st = __uint_as_float(cur.y);
                    dist = __fadd_rn(dist, st);              // distance += sign * t, ascending triangle order
                    ssum += (0.0f < st) - (st < 0.0f);
                    ++hits; ++i;
                    if (i < n) cur = entry(i);
                }
            }
            L = attenuate(L, M.mu, dist, ssum, a.length_to_cm);
            if (m == 0) { dist0 = dist; ssum0 = ssum; hits0 = hits; }
        }
    }
    return { L, dist0, ssum0, hits0 };
}

__global__ void __launch_bounds__(256, SRT_RESOLVE_MIN_CTAS) resolve_hits_kernel(const __grid_constant__ RenderArgs a)
{
    __shared__ float s_sort[8][kSortOverflowCap];
    __shared__ PixelOut s_none;                              // outputs of a pixel no triangle touches (~70 % of a frame): computed once
    if (threadIdx.x == 0) s_none = resolve_list(a, kNoHit);
    __syncthreads();
    const uint32_t W = a.cam.W;
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t by = blockIdx.y >> 4, lrow = blockIdx.y & 15u;
    const uint32_t row = a.row_begin + (by * a.blk_stride + a.blk_phase) * 16 + lrow;
    const uint32_t orow = a.compact ? by * 16 + lrow : row;
    const bool valid = col < W && row < a.row_end;          // no early return: the overflow path below is warp-wide
    const size_t p = (size_t)orow * W + (valid ? col : 0);
    float L = a.incident;
    float dist0 = 0.0f; int ssum0 = 0; uint32_t hits0 = 0;   // mode 1: z-buffer value / triangle index
    if (a.mode == 1) { dist0 = INFINITY; ssum0 = -1; }
    // a candidate list overflowed: no hit list can be trusted, every pixel is recomputed below
    const bool overflow = a.cand_count[0] > a.cand_cap || a.cand_count[1] > a.pix_cap || (a.big_count && *a.big_count > a.big_cap);
    if (valid) {
        const uint32_t head = a.hit_head[p];
        if (head != kNoHit) a.hit_head[p] = kNoHit;         // ready for the next render
        if (!overflow) {
            const PixelOut r = head == kNoHit ? s_none : resolve_list(a, head);
            L = r.L; dist0 = r.dist; ssum0 = r.ssum; hits0 = r.hits;
        }
    }
    // The candidate lists overflowed: the WARP recomputes each of its pixels from scratch with the exact sequence --
    // 32 triangles at a time (one per lane), hits folded in ascending lane (= triangle) order.
    const uint32_t lane = threadIdx.x & 31u;
    unsigned todo = __ballot_sync(0xffffffffu, valid && overflow);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t prow = __shfl_sync(0xffffffffu, row, src), pcol = __shfl_sync(0xffffffffu, col, src);
        const RayDirs r = make_ray(a.cam, prow, pcol);
        float Lc = a.incident, d0 = a.mode == 1 ? INFINITY : 0.0f; int s0 = a.mode == 1 ? -1 : 0; uint32_t h0 = 0;
        float* ts = s_sort[threadIdx.x >> 5];               // sort mode: the ray's distances, in the warp's scratch
        uint32_t n_sort = 0;
        for (int m = 0; m < a.n_meshes; ++m) {
            const DevMesh M = a.meshes[m];
            float dist = 0.0f; int ssum = 0; uint32_t hits = 0;
            if (M.contributes) {
#pragma unroll 1
                for (uint32_t base = 0; base < M.n; base += 32) {
                    const uint32_t j = base + lane;
                    float st = 0.0f; bool got = false;
                    if (j < M.n) {
                        const float4 A = __ldg(M.exact + 4 * (size_t)j), B = __ldg(M.exact + 4 * (size_t)j + 1);
                        const float4 C = __ldg(M.exact + 4 * (size_t)j + 2), Dn = __ldg(M.exact + 4 * (size_t)j + 3);
                        float t;
                        if (exact_intersect(r, A, B, C, Dn.x, t)) {
                            if (a.mode == 1) { st = t; got = true; }
                            else if (a.mode == 2) { st = t; got = (double)t > 0.0000001; }
                            else if ((double)t > 0.0000001) {
                                const float dn = dot_rn(r.ax, r.ay, r.az, Dn.y, Dn.z, Dn.w);
                                st = __fmul_rn((float)((0.0f < dn) - (dn < 0.0f)), t);
                                got = true;
                            }
                        }
                    }
                    unsigned hm = __ballot_sync(0xffffffffu, got);
                    while (hm) {
                        const int l = __ffs(hm) - 1;
                        hm &= hm - 1;
                        const float v = __shfl_sync(0xffffffffu, st, l);
                        if (a.mode == 1) { if (d0 > v) { d0 = v; s0 = (int)(M.first + base + (uint32_t)l); } }
                        else if (a.mode == 2) { if (n_sort < kSortOverflowCap && lane == 0) ts[n_sort] = v; ++n_sort; }
                        else { dist = __fadd_rn(dist, v); ssum += (0.0f < v) - (v < 0.0f); ++hits; }
                    }
                }
            }
            if (a.mode != 0) continue;
            Lc = attenuate(Lc, M.mu, dist, ssum, a.length_to_cm);
            if (m == 0) { d0 = dist; s0 = ssum; h0 = hits; }
        }
        if (a.mode == 2) {
            __syncwarp();
            if ((int)lane == src) {
                // more intersections on one ray than the scratch holds: not representable here (NaN, never a wrong number)
                if (n_sort > kSortOverflowCap) { L = NAN; dist0 = NAN; }
                else L = sorted_xray_value(ts, n_sort, a.incident, a.meshes[0].mu, a.length_to_cm, dist0);
                ssum0 = 0; hits0 = n_sort;
            }
            __syncwarp();
            continue;
        }
        if ((int)lane == src) { L = Lc; dist0 = d0; ssum0 = s0; hits0 = h0; }
    }
    if (!valid) return;
    if (a.mode != 1) store_L(a, a.L_image ? (size_t)row * W + col : p, row, L);
    if (a.dist) a.dist[p] = dist0;
    if (a.ssum) a.ssum[p] = ssum0;
    if (a.hits) a.hits[p] = hits0;
}

// ------------------------------------------------------------------ culled kernels
// SRT_KERNEL_CULLED: the same conservative strip tests, applied hierarchically.  A pixel p inside an axis-aligned box
// of pixel centres (centre (xc, yc), half extents hx, hy) has |Y(p)| >= |Y(c)| - r_Y and |D(p)| <= |D(c)| + r_D with
// r = |beta| hx + |gamma| hy, so "no pixel of the box can pass strip i" follows from
//     |Y_i(c)| - |D(c)| > r_Yi + r_D + 2 m_i
// (m_i already bounds the float evaluation error at every point of the frame -- twice: once for the pixel, once for
// the box centre, which is always inside the frame because boxes are clamped to it; the factor kBoxInflate covers the
// roundings of the radius sum itself).  Three levels use it: bin (256 x 128 pixels, one thread per triangle, ordered
// compaction into per-bin index lists), CTA tile (64 x 16, one thread per listed triangle, ordered compaction into
// shared memory) and the thread's run of 8 pixels; survivors of all three take the per-pixel tests and the exact
// reference sequence exactly like the slow path of the filtered kernel -- in ascending triangle order per pixel, so
// the float sums are bit-identical.  Every (pixel, triangle) pair is still decided by the same proven-conservative
// filter or the exact test; none is skipped on a heuristic.
// Bin pass: one CTA per (bin, mesh); thread t of iteration i tests triangle 1024 i + t; survivors are appended in
// ascending triangle order (ballot + warp-count scan).
constexpr uint32_t kBinThreads = 1024;
__global__ void __launch_bounds__(kBinThreads) bin_kernel(const __grid_constant__ RenderArgs a)
{
    __shared__ uint32_t warp_cnt[kBinThreads / 32];
    const int m = (int)blockIdx.z;
    const DevMesh M = a.meshes[m];
    const uint32_t bin = blockIdx.y * a.bins_x + blockIdx.x;
    uint32_t* cnt_out = a.bin_counts + (size_t)bin * a.n_meshes + m;
    if (!M.contributes || M.n == 0) { if (threadIdx.x == 0) *cnt_out = 0; return; }
    const uint32_t c0 = blockIdx.x * kBinCols, c1 = min(a.cam.W, c0 + kBinCols) - 1;
    const uint32_t b0 = blockIdx.y * a.bin_blocks, b1 = min(a.n_blocks, b0 + a.bin_blocks) - 1;
    const uint32_t r0 = a.row_begin + (b0 * a.blk_stride + a.blk_phase) * 16;
    const uint32_t r1 = min(a.row_end, a.row_begin + (b1 * a.blk_stride + a.blk_phase) * 16 + 16) - 1;
    const float xc = 0.5f * (float)(c0 + c1) + a.cam.xc, hx = 0.5f * (float)(c1 - c0);
    const float yc = 0.5f * (float)(r0 + r1) + a.cam.yc, hy = 0.5f * (float)(r1 - r0);
    // the mesh's share of the bin's list: proportional to its triangle count; survivors beyond it are counted but not
    // stored, and the render pass then walks the whole mesh for this bin (count > capacity)
    const uint32_t off_m = (uint32_t)((uint64_t)M.first * a.bin_cap / a.total_tris);
    const uint32_t cap_m = (uint32_t)((uint64_t)(M.first + M.n) * a.bin_cap / a.total_tris) - off_m;
    uint32_t* out = a.bin_entries + (size_t)bin * a.bin_cap + off_m;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = 0;
    uint32_t j = threadIdx.x;
    float4 fa, fb, ga, gb;
    auto load_rec = [&](uint32_t i) {
        const float4* q0 = M.filt0 + 3 * (size_t)i; const float4* q1 = M.filt1 + 2 * (size_t)i;
        fa = __ldg(q0); fb = __ldg(q0 + 1); ga = __ldg(q1); gb = __ldg(q1 + 1);
    };
    if (j < M.n) load_rec(j);
    for (uint32_t j0 = 0; j0 < M.n; j0 += kBinThreads) {
        const bool keep = j < M.n && box_may_hit(fa, fb, ga, gb, xc, yc, hx, hy);
        const uint32_t jn = j + kBinThreads;                   // prefetch the next iteration's record
        if (jn < M.n) load_rec(jn);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t v = warp_cnt[lane];                           // kBinThreads / 32 == 32 warps: one count per lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += u; }
        const uint32_t total = __shfl_sync(0xffffffffu, v, 31);
        const uint32_t before = warp ? __shfl_sync(0xffffffffu, v, (int)warp - 1) : 0u;
        if (keep) { const uint32_t pos = base + before + __popc(bal & ((1u << lane) - 1u)); if (pos < cap_m) out[pos] = j; }
        base += total;
        j = jn;
        __syncthreads();                                       // warp_cnt is rewritten next iteration
    }
    if (threadIdx.x == 0) *cnt_out = base;
}

struct CulledArgs {
    const float4* exact;
    uint32_t first; int mode;
    float y, x0, run_xc, run_hx; uint32_t col_end;      // x0: centred x of the run's first pixel; col_end = min(W - col0, 8)
    uint32_t wrow0, wcol0;                              // first row / column of the warp's 32 x 8 pixel tile
    uint8_t* queue;                                     // this warp's candidate queue (256 entries)
    float* acc;                                         // this warp's accumulators: [3][256], slot = k * 32 + lane
};

// One listed triangle for one WARP (called convergently): every lane applies the run-level box test and the per-pixel
// strip tests to its run of 8 pixels; the surviving (pixel, triangle) pairs of the whole warp are queued in shared
// memory and the exact reference sequence is run on them with all lanes busy (a triangle typically covers most of
// the pixels of two or three lanes: lane-serial execution would leave the warp 90 % idle).  The executing lane
// regenerates the pixel's ray and folds the result into the pixel's accumulators in shared memory: a triangle
// contributes at most once per pixel and triangles are processed strictly in ascending order with a warp barrier in
// between, so every per-pixel float sum keeps the reference's order.
__device__ __noinline__ void culled_process(const DevCamera& cam, const CulledArgs& s, const float4* rec, uint32_t jj,
                                            unsigned long long& n_exact)
{
    constexpr int P = 8;
    const uint32_t lane = threadIdx.x & 31u;
    const float4 fa = rec[0], fb = rec[1], ga = rec[2], gb = rec[3];                        // shared memory
    uint32_t mask = 0;
    if (s.col_end && box_may_hit(fa, fb, ga, gb, s.run_xc, s.y, s.run_hx, 0.0f)) {
        const float bs = fmaf(fa.y, s.y, fa.z), bd = fmaf(fb.x, s.y, fb.y);
        const float b2 = fmaf(ga.y, s.y, ga.z), b3 = fmaf(gb.y, s.y, gb.z);
        const float m1 = fb.z, m2 = ga.w, m3 = gb.w;
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const float xk = s.x0 + (float)k;                              // exact: (half-)integers far below 2^23
            const float Dd = fabsf(fmaf(fa.w, xk, bd));
            const float S = fmaf(fa.x, xk, bs), X2 = fmaf(ga.x, xk, b2), X3 = fmaf(gb.x, xk, b3);
            const bool reject = (fabsf(S) > Dd + m1) || (fabsf(X2) > Dd + m2) || (fabsf(X3) > Dd + m3);   // NaN survives
            if (!reject) mask |= 1u << k;
        }
        mask &= (1u << s.col_end) - 1u;                                    // pixels beyond the last column
    }
    if (!__any_sync(0xffffffffu, mask != 0)) return;
    // queue positions: exclusive prefix of the per-lane candidate counts
    const uint32_t mine = __popc(mask);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    {
        uint32_t mm = mask, q = incl - mine;
        while (mm) { const int k = __ffs(mm) - 1; mm &= mm - 1; s.queue[q++] = (uint8_t)(((uint32_t)k << 5) | lane); }
    }
    __syncwarp();
    n_exact += mine;
    const float4 A = rec[4], B = rec[5], C = rec[6], Dn = rec[7];          // ExactTri, staged with the filter record
    for (uint32_t q = lane; q < total; q += 32) {
        const uint32_t e = s.queue[q], owner = e & 31u, k = e >> 5;        // e is also the accumulator slot k * 32 + owner
        const RayDirs r = make_ray(cam, s.wrow0 + (owner >> 2), s.wcol0 + (owner & 3u) * P + k);
        float tt;
        if (exact_intersect(r, A, B, C, Dn.x, tt)) {
            float* acc = s.acc + e;
            if (s.mode != 0) {                                             // closest hit, strict > (main-omp.cxx:656-662)
                if (acc[0] > tt) { acc[0] = tt; acc[256] = __int_as_float((int)(s.first + jj)); }
            } else if ((double)tt > 0.0000001) {                           // lbuffer.cxx:788
                const float dn = dot_rn(r.ax, r.ay, r.az, Dn.y, Dn.z, Dn.w);
                const int sign = (0.0f < dn) - (dn < 0.0f);
                acc[0] = __fadd_rn(acc[0], __fmul_rn((float)sign, tt));    // distance += sign * t
                acc[256] = __int_as_float(__float_as_int(acc[256]) + sign);
                acc[512] = __int_as_float(__float_as_int(acc[512]) + 1);
            }
        }
    }
    __syncwarp();                                                          // queue and accumulators: next triangle
}

// Render pass: CTA = 64 x 16 pixels, thread = run of 8 pixels of one row (as in the filtered kernel), but the four
// warps (32 x 8 pixels each) work independently: each walks the bin's list 32 entries at a time (one listed triangle
// per lane against the warp's tile), compacts the survivors into its own shared-memory segment and processes them;
// there is no CTA-wide barrier, so a warp over empty space does not wait for a warp over the object.
#ifndef SRT_CULLED_MIN_CTAS
#define SRT_CULLED_MIN_CTAS 4
#endif
#ifndef SRT_CULLED_CTA_LIST
#define SRT_CULLED_CTA_LIST 1
#endif
constexpr uint32_t kCtaList = 1024;
__global__ void __launch_bounds__(128, SRT_CULLED_MIN_CTAS) lbuffer_culled_kernel(const __grid_constant__ RenderArgs a)
{
    constexpr int P = 8;
    __shared__ float4 s_rec[4][32][8];          // per-warp survivors: filter record (4 x float4) + ExactTri (4 x float4)
    __shared__ uint32_t s_idx[4][32];
    __shared__ float s_acc[4][3 * 256];         // per-warp accumulators: dist | ssum | hits (z | id in closest-hit mode)
    __shared__ uint8_t s_queue[4][256];
#if SRT_CULLED_CTA_LIST
    __shared__ uint32_t s_list[kCtaList];       // triangles of the bin that may touch the CTA tile, ascending
    __shared__ uint32_t s_wcnt[4];
#endif

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CtaCoord cta = cta_coord();
    const uint32_t wcol0 = (cta.x * 8 + (warp & 1) * 4) * P;
    const uint32_t col0 = wcol0 + (lane & 3) * P;
    const uint32_t row0 = a.row_begin + (cta.y * a.blk_stride + a.blk_phase) * 16;
    const uint32_t wrow0 = row0 + (warp >> 1) * 8;
    const uint32_t row = wrow0 + (lane >> 2);
    const uint32_t orow = a.compact ? cta.y * 16 + (warp >> 1) * 8 + (lane >> 2) : row;
    const uint32_t W = a.cam.W;
    const bool row_ok = row < a.row_end;
    const bool active = row_ok && col0 < W;
    const bool warp_on = wrow0 < a.row_end && wcol0 < W;       // warp-uniform: the warp's tile has pixels

    // warp tile box and this thread's run box, clamped to the frame
    const uint32_t tc1 = min(W, wcol0 + 32) - 1, tr1 = min(a.row_end, wrow0 + 8) - 1;
    const float txc = 0.5f * (float)(wcol0 + tc1) + a.cam.xc, thx = 0.5f * (float)(tc1 - wcol0);
    const float tyc = 0.5f * (float)(wrow0 + tr1) + a.cam.yc, thy = 0.5f * (float)(tr1 - wrow0);
    const uint32_t rc1 = min(W, col0 + P) - 1;
#if SRT_CULLED_CTA_LIST
    // CTA tile box (64 x 16 pixels), clamped to the frame
    const uint32_t cc0 = cta.x * 64, cc1 = min(W, cc0 + 64) - 1, cr1 = min(a.row_end, row0 + 16) - 1;
    const float cxc = 0.5f * (float)(cc0 + cc1) + a.cam.xc, chx = 0.5f * (float)(cc1 - cc0);
    const float cyc = 0.5f * (float)(row0 + cr1) + a.cam.yc, chy = 0.5f * (float)(cr1 - row0);
#endif

    float Lval[P];
#pragma unroll
    for (int k = 0; k < P; ++k) Lval[k] = a.incident;
    unsigned long long n_exact = 0;
    float* acc = s_acc[warp];
    if (a.mode != 0) {
#pragma unroll
        for (int k = 0; k < P; ++k) { acc[k * 32 + lane] = INFINITY; acc[256 + k * 32 + lane] = __int_as_float(-1); }
    }

    const uint32_t bin = (cta.y / a.bin_blocks) * a.bins_x + cta.x / (kBinCols / 64);
    for (int m = 0; m < a.n_meshes; ++m) {
        const DevMesh M = a.meshes[m];
        if (a.mode == 0) {
#pragma unroll
            for (int k = 0; k < P; ++k) { acc[k * 32 + lane] = 0.0f; acc[256 + k * 32 + lane] = 0.0f; acc[512 + k * 32 + lane] = 0.0f; }
        }
        __syncwarp();
        if (M.contributes && M.n > 0) {
            CulledArgs sa;
            sa.exact = M.exact; sa.first = M.first; sa.mode = a.mode;
            sa.y = (float)row + a.cam.yc; sa.x0 = (float)col0 + a.cam.xc;
            sa.queue = s_queue[warp]; sa.acc = acc; sa.wrow0 = wrow0; sa.wcol0 = wcol0;
            sa.col_end = active ? rc1 - col0 + 1 : 0u;
            sa.run_xc = 0.5f * (float)(col0 + rc1) + a.cam.xc; sa.run_hx = active ? 0.5f * (float)(rc1 - col0) : 0.0f;
            const uint32_t off_m = (uint32_t)((uint64_t)M.first * a.bin_cap / a.total_tris);
            const uint32_t cap_m = (uint32_t)((uint64_t)(M.first + M.n) * a.bin_cap / a.total_tris) - off_m;
            const uint32_t listed = __ldg(a.bin_counts + (size_t)bin * a.n_meshes + m);
            const bool whole_mesh = listed > cap_m;                  // the bin's list overflowed: every triangle is a candidate
            const uint32_t cnt = whole_mesh ? M.n : listed;
            const uint32_t* list = a.bin_entries + (size_t)bin * a.bin_cap + off_m;
#if SRT_CULLED_CTA_LIST
            // Phase A (whole CTA, equal work per warp): cull the bin's list against the CTA tile, 128 entries per step, into
            // an ordered index list in shared memory.  Phase B (warps independent): each warp culls that short list against
            // its own 32 x 8 tile and processes the survivors.  A list longer than kCtaList is handled in rounds.
            uint32_t idx = 0; float4 fa, fb, ga, gb;
            if (tid < cnt) {
                idx = (whole_mesh ? tid : __ldg(list + tid));
                const float4* q0 = M.filt0 + 3 * (size_t)idx; const float4* q1 = M.filt1 + 2 * (size_t)idx;
                fa = __ldg(q0); fb = __ldg(q0 + 1); ga = __ldg(q1); gb = __ldg(q1 + 1);
            }
            uint32_t i0 = 0;
            while (i0 < cnt) {                                                 // rounds (CTA-uniform)
                uint32_t n_list = 0;
                while (i0 < cnt && n_list + 128 <= kCtaList) {
                    const bool keep = i0 + tid < cnt && box_may_hit(fa, fb, ga, gb, cxc, cyc, chx, chy);
                    const uint32_t my_idx = idx;
                    const uint32_t in = i0 + 128 + tid;                       // next step's records in flight
                    if (in < cnt) {
                        idx = (whole_mesh ? in : __ldg(list + in));
                        const float4* q0 = M.filt0 + 3 * (size_t)idx; const float4* q1 = M.filt1 + 2 * (size_t)idx;
                        fa = __ldg(q0); fb = __ldg(q0 + 1); ga = __ldg(q1); gb = __ldg(q1 + 1);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (lane == 0) s_wcnt[warp] = __popc(bal);
                    __syncthreads();
                    uint32_t off = 0, tot = 0;
#pragma unroll
                    for (uint32_t w = 0; w < 4; ++w) { const uint32_t c = s_wcnt[w]; off += w < warp ? c : 0u; tot += c; }
                    if (keep) s_list[n_list + off + __popc(bal & ((1u << lane) - 1u))] = my_idx;
                    n_list += tot; i0 += 128;
                    __syncthreads();                                           // s_wcnt is rewritten; s_list is visible
                }
                if (warp_on) {
                    for (uint32_t j0 = 0; j0 < n_list; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        bool keep = false; uint32_t id2 = 0;
                        float4 ra, rb, rc, rd;
                        if (j < n_list) {
                            id2 = s_list[j];
                            const float4* q0 = M.filt0 + 3 * (size_t)id2; const float4* q1 = M.filt1 + 2 * (size_t)id2;     // L1 / L2 hit: phase A has just read it
                            ra = __ldg(q0); rb = __ldg(q0 + 1); rc = __ldg(q1); rd = __ldg(q1 + 1);
                            keep = box_may_hit(ra, rb, rc, rd, txc, tyc, thx, thy);
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, keep);
                        if (bal == 0) continue;
                        if (keep) {
                            const uint32_t pos = __popc(bal & ((1u << lane) - 1u));
                            const float4* eq = M.exact + 4 * (size_t)id2;    // the exact constants travel with the survivor
                            const float4 e0 = __ldg(eq), e1 = __ldg(eq + 1), e2 = __ldg(eq + 2), e3 = __ldg(eq + 3);
                            float4* dst = &s_rec[warp][pos][0];
                            dst[0] = ra; dst[1] = rb; dst[2] = rc; dst[3] = rd; dst[4] = e0; dst[5] = e1; dst[6] = e2; dst[7] = e3;
                            s_idx[warp][pos] = id2;
                        }
                        __syncwarp();
                        const uint32_t n_w = __popc(bal);
#pragma unroll 1
                        for (uint32_t q = 0; q < n_w; ++q)
                            culled_process(a.cam, sa, &s_rec[warp][q][0], s_idx[warp][q], n_exact);
                        __syncwarp();                                          // the segment is rewritten by the next chunk
                    }
                }
                __syncthreads();                                               // every warp is done with s_list
            }
#else
            if (warp_on) {
            // software pipeline: the next 32 entries' records are in flight while the current ones are processed
            uint32_t idx = 0; float4 fa, fb, ga, gb;
            if (lane < cnt) {
                idx = (whole_mesh ? lane : __ldg(list + lane));
                const float4* q0 = M.filt0 + 3 * (size_t)idx; const float4* q1 = M.filt1 + 2 * (size_t)idx;
                fa = __ldg(q0); fb = __ldg(q0 + 1); ga = __ldg(q1); gb = __ldg(q1 + 1);
            }
            for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
                const bool keep = i0 + lane < cnt && box_may_hit(fa, fb, ga, gb, txc, tyc, thx, thy);
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const uint32_t pos = __popc(bal & ((1u << lane) - 1u));
                    const float4* eq = M.exact + 4 * (size_t)idx;          // the exact constants travel with the survivor
                    const float4 e0 = __ldg(eq), e1 = __ldg(eq + 1), e2 = __ldg(eq + 2), e3 = __ldg(eq + 3);
                    float4* dst = &s_rec[warp][pos][0];
                    dst[0] = fa; dst[1] = fb; dst[2] = ga; dst[3] = gb; dst[4] = e0; dst[5] = e1; dst[6] = e2; dst[7] = e3;
                    s_idx[warp][pos] = idx;
                }
                const uint32_t in = i0 + 32 + lane;
                if (in < cnt) {
                    idx = (whole_mesh ? in : __ldg(list + in));
                    const float4* q0 = M.filt0 + 3 * (size_t)idx; const float4* q1 = M.filt1 + 2 * (size_t)idx;
                    fa = __ldg(q0); fb = __ldg(q0 + 1); ga = __ldg(q1); gb = __ldg(q1 + 1);
                }
                if (bal == 0) continue;
                __syncwarp();
                const uint32_t n_w = __popc(bal);
#pragma unroll 1
                for (uint32_t q = 0; q < n_w; ++q)
                    culled_process(a.cam, sa, &s_rec[warp][q][0], s_idx[warp][q], n_exact);
                __syncwarp();                                  // the segment is rewritten by the next chunk
            }
            }
#endif
        }
        if (a.mode != 0) continue;
        const bool out_ok = m == 0 && row_ok;
        const bool plain = isfinite(M.mu) && isfinite(a.length_to_cm);     // else 0 * inf = NaN must reach exp()
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const float dist = acc[k * 32 + lane];
            const int ssum = __float_as_int(acc[256 + k * 32 + lane]);
            // exp(-(mu * (0 * 0.1))) == 1 exactly: pixels the mesh does not touch keep their value bit for bit
            if (dist != 0.0f || ssum != 0 || !plain) Lval[k] = attenuate(Lval[k], M.mu, dist, ssum, a.length_to_cm);
            if (out_ok && col0 + k < W) {
                const size_t p = (size_t)orow * W + col0 + k;
                if (a.dist) a.dist[p] = dist;
                if (a.ssum) a.ssum[p] = ssum;
                if (a.hits) a.hits[p] = (uint32_t)__float_as_int(acc[512 + k * 32 + lane]);
            }
        }
    }
    if (a.mode != 0) {
        if (row_ok) {
#pragma unroll
            for (int k = 0; k < P; ++k)
                if (col0 + k < W) {
                    a.dist[(size_t)orow * W + col0 + k] = acc[k * 32 + lane];
                    a.ssum[(size_t)orow * W + col0 + k] = __float_as_int(acc[256 + k * 32 + lane]);
                }
        }
    } else if (row_ok) {
        const size_t out = (size_t)(a.L_image ? row : orow) * W + col0;
        if (a.L_vec4 && col0 + P <= W) {
#pragma unroll
            for (int k = 0; k < P; k += 4)
                store_L4(a, out + k, row, make_float4(Lval[k], Lval[k + 1], Lval[k + 2], Lval[k + 3]));
        } else {
#pragma unroll
            for (int k = 0; k < P; ++k)
                if (col0 + k < W) store_L(a, out + k, row, Lval[k]);
        }
    }
    if (a.counters && n_exact) atomicAdd(a.counters, n_exact);
}

// ------------------------------------------------------------------ shadow grid
__device__ __forceinline__ uint32_t float_to_ordered_key(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ int mesh_of(const ShadeMesh* meshes, int n_meshes, uint32_t g)
{
    int m = 0;
    while (m + 1 < n_meshes && g >= meshes[m + 1].first) ++m;
    return m;
}

__global__ void __launch_bounds__(256) scene_bbox_kernel(const ShadeMesh* meshes, int n_meshes, uint32_t total, uint32_t* out6)
{
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        const ShadeMesh M = meshes[mesh_of(meshes, n_meshes, g)];
        const float* v = M.vertices + 9 * (size_t)(g - M.first);
        for (int k = 0; k < 9; ++k) { const float x = __ldg(v + k); if (x == x) { lo[k % 3] = fminf(lo[k % 3], x); hi[k % 3] = fmaxf(hi[k % 3], x); } }
    }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(out6 + k, float_to_ordered_key(lo[k]));
            atomicMax(out6 + 3 + k, float_to_ordered_key(hi[k]));
        }
    }
}

// central projection of X from the light: (a, b) = (q.u, q.v) / q.w; false when X is too close to the plane through
// the light perpendicular to w
__device__ __forceinline__ bool light_project(const ShadowGrid& g, const float* L, float x, float y, float z, float& pa, float& pb, float& depth)
{
    const float qx = x - L[0], qy = y - L[1], qz = z - L[2];
    depth = fmaf(qx, g.w[0], fmaf(qy, g.w[1], qz * g.w[2]));
    const float qn = fabsf(qx) + fabsf(qy) + fabsf(qz);
    if (!(fabsf(depth) > 1e-3f * qn)) return false;                      // also NaN
    const float inv = 1.0f / depth;
    pa = fmaf(qx, g.u[0], fmaf(qy, g.u[1], qz * g.u[2])) * inv;
    pb = fmaf(qx, g.v[0], fmaf(qy, g.v[1], qz * g.v[2])) * inv;
    return true;
}

template <bool FILL>
__global__ void __launch_bounds__(256) shadow_bin_kernel(const __grid_constant__ ShadeArgs a)
{
    const ShadowGrid& g = a.grid;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.total_tris) return;
    const ShadeMesh M = a.meshes[mesh_of(a.meshes, a.n_meshes, t)];
    const float* v = M.vertices + 9 * (size_t)(t - M.first);
    float pa[3], pb[3], dp[3];
    bool ok = true;
    for (int k = 0; k < 3; ++k) ok = light_project(g, a.light_position, __ldg(v + 3 * k), __ldg(v + 3 * k + 1), __ldg(v + 3 * k + 2), pa[k], pb[k], dp[k]) && ok;
    ok = ok && ((dp[0] > 0.0f) == (dp[1] > 0.0f)) && ((dp[0] > 0.0f) == (dp[2] > 0.0f));     // all on one side of the light
    int ia0 = 0, ia1 = -1, ib0 = 0, ib1 = -1;
    if (ok) {
        const float amin = fminf(pa[0], fminf(pa[1], pa[2])), amax = fmaxf(pa[0], fmaxf(pa[1], pa[2]));
        const float bmin = fminf(pb[0], fminf(pb[1], pb[2])), bmax = fmaxf(pb[0], fmaxf(pb[1], pb[2]));
        // cell range widened by one cell on every side, clamped to the grid (queries outside the grid use brute force)
        const float fa0 = floorf((amin - g.a0) * g.inv_da) - 1.0f, fa1 = floorf((amax - g.a0) * g.inv_da) + 1.0f;
        const float fb0 = floorf((bmin - g.b0) * g.inv_db) - 1.0f, fb1 = floorf((bmax - g.b0) * g.inv_db) + 1.0f;
        ok = fa0 == fa0 && fa1 == fa1 && fb0 == fb0 && fb1 == fb1;       // NaN: every ray tests it
        if (ok) {
            const float Gf = (float)g.G;
            ia0 = (int)fminf(fmaxf(fa0, 0.0f), Gf); ia1 = (int)fminf(fmaxf(fa1, -1.0f), Gf - 1.0f);
            ib0 = (int)fminf(fmaxf(fb0, 0.0f), Gf); ib1 = (int)fminf(fmaxf(fb1, -1.0f), Gf - 1.0f);
        }
    }
    const long long cells = ok ? (long long)max(0, ia1 - ia0 + 1) * (long long)max(0, ib1 - ib0 + 1) : 0;
    if (!ok || cells > (long long)kShadowMaxCells) {
        if (FILL) g.global_tris[atomicAdd(g.n_global, 1u)] = t;
        return;
    }
    for (int ib = ib0; ib <= ib1; ++ib)
        for (int ia = ia0; ia <= ia1; ++ia) {
            const uint32_t cell = (uint32_t)ib * g.G + (uint32_t)ia;
            if (FILL) g.cell_tris[atomicAdd(g.cell_cursor + cell, 1u)] = t;
            else atomicAdd(g.cell_off + cell, 1u);
        }
}

// in-place exclusive scan of n values by one 1024-thread CTA (n ~ 2.6e5 cells: ~20 us)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(uint32_t* data, uint32_t n)
{
    __shared__ uint32_t warp_sum[32];
    const uint32_t per = (n + 1023) / 1024, lo = threadIdx.x * per, hi = min(n, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += data[i];
    uint32_t incl = sum;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = warp_sum[lane];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += u; }
        warp_sum[lane] = v;
    }
    __syncthreads();
    uint32_t run = incl - sum + (warp ? warp_sum[warp - 1] : 0u);
    for (uint32_t i = lo; i < hi; ++i) { const uint32_t c = data[i]; data[i] = run; run += c; }
}

// ------------------------------------------------------------------ shade kernel
// Shaded variant, second pass (src/main-omp.cxx:668-794): Phong shading of the closest hit, shadow ray against
// every other triangle, area-barycentric texture lookup, clamp + truncation to 8 bits.  One thread per pixel.
// Shadow rays have per-pixel origins, so the common-origin pre-pack does not apply: the test below is the
// reference's full sequence on the raw vertices.
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3sub(V3 a, V3 b) { return { __fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z) }; }
__device__ __forceinline__ V3 v3add(V3 a, V3 b) { return { __fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z) }; }
__device__ __forceinline__ V3 v3mulf(V3 a, float t) { return { __fmul_rn(a.x, t), __fmul_rn(a.y, t), __fmul_rn(a.z, t) }; }
__device__ __forceinline__ V3 v3mul(V3 a, V3 b) { return { __fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z) }; }
__device__ __forceinline__ V3 v3muld(V3 a, double t)       // Vec3::operator*(double): double product, narrowed
{ return { (float)__dmul_rn((double)a.x, t), (float)__dmul_rn((double)a.y, t), (float)__dmul_rn((double)a.z, t) }; }
__device__ __forceinline__ float v3dot(V3 a, V3 b) { return dot_rn(a.x, a.y, a.z, b.x, b.y, b.z); }
__device__ __forceinline__ V3 v3cross(V3 a, V3 b)
{
    return { __fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
             __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)) };
}
__device__ __forceinline__ V3 v3normalise(V3 a)
{ const float l = __fsqrt_rn(v3dot(a, a)); return { __fdiv_rn(a.x, l), __fdiv_rn(a.y, l), __fdiv_rn(a.z, l) }; }
__device__ __forceinline__ V3 ld_v3(const float* p) { return { __ldg(p), __ldg(p + 1), __ldg(p + 2) }; }

// Ray::intersect (src/Ray.cxx:72-125) for an arbitrary origin
__device__ __forceinline__ bool intersect_generic(V3 o, V3 d, V3 p1, V3 p2, V3 p3, float& t)
{
    const V3 e1 = v3sub(p2, p1), e2 = v3sub(p3, p1);
    const V3 pvec = v3cross(d, e2);
    const float det = v3dot(e1, pvec);
    if (det == 0.0f) return false;
    const float inv = __fdiv_rn(1.0f, det);
    const V3 tvec = v3sub(o, p1);
    const float u = __fmul_rn(v3dot(tvec, pvec), inv);
    if (u < 0.0f || u > 1.0f) return false;
    const V3 qvec = v3cross(tvec, e1);
    const float v = __fmul_rn(v3dot(d, qvec), inv);
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    t = __fmul_rn(v3dot(e2, qvec), inv);
    return true;
}

// Triangle::getArea (include/Triangle.inl:206-216): float cross components, double pow/sqrt, narrowed
__device__ __forceinline__ float tri_area(V3 a, V3 b, V3 c)
{
    const V3 ab = v3sub(b, a), ac = v3sub(c, a);
    const double x = (double)__fsub_rn(__fmul_rn(ab.y, ac.z), __fmul_rn(ab.z, ac.y));
    const double y = (double)__fsub_rn(__fmul_rn(ab.z, ac.x), __fmul_rn(ab.x, ac.z));
    const double z = (double)__fsub_rn(__fmul_rn(ab.x, ac.y), __fmul_rn(ab.y, ac.x));
    return (float)__dmul_rn(0.5, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z))));
}

__device__ __forceinline__ uint8_t clamp_trunc(float c) { return c < 0.0f ? 0 : (c > 255.0f ? 255 : (uint8_t)(int)c); }
__device__ __forceinline__ uint8_t clamp_trunc(double c) { return c < 0.0 ? 0 : (c > 255.0 ? 255 : (uint8_t)(int)c); }

__global__ void __launch_bounds__(128) shade_kernel(const __grid_constant__ ShadeArgs a)
{
    const uint32_t col = blockIdx.x * 8 + (threadIdx.x & 7);
    const uint32_t row = a.row_begin + blockIdx.y * 16 + (threadIdx.x >> 3);
    const uint32_t W = a.cam.W;
    if (col >= W || row >= a.row_end) return;
    const size_t pixel = (size_t)row * W + col;
    const int id = a.tri_id[pixel];
    uint8_t* out = a.rgb + 3 * pixel;
    if (id < 0) { out[0] = a.background[0]; out[1] = a.background[1]; out[2] = a.background[2]; if (a.shadow) a.shadow[pixel] = 0; return; }
    int hm = 0;
    while (hm + 1 < a.n_meshes && (uint32_t)id >= a.meshes[hm + 1].first) ++hm;
    const ShadeMesh M = a.meshes[hm];
    const uint32_t local = (uint32_t)id - M.first;
    const RayDirs r = make_ray(a.cam, row, col);
    const V3 origin = { a.cam.ox, a.cam.oy, a.cam.oz }, rdir = { r.bx, r.by, r.bz }, dirA = { r.ax, r.ay, r.az };
    const float t = a.z[pixel];
    const V3 P = v3add(origin, v3mulf(rdir, t));                                       // point_hit (:671)
    const float4 En = __ldg(M.exact + 4 * (size_t)local + 3);
    const V3 N = { En.y, En.z, En.w };
    // ---- applyShading (:407-434)
    const V3 Lc = { a.light_colour[0], a.light_colour[1], a.light_colour[2] };
    const V3 Lp = { a.light_position[0], a.light_position[1], a.light_position[2] };
    const V3 Ka = { M.ka[0], M.ka[1], M.ka[2] }, Kd = { M.kd[0], M.kd[1], M.kd[2] }, Ks = { M.ks[0], M.ks[1], M.ks[2] };
    const V3 ambient = v3mul(Lc, Ka);
    const V3 lightDir = v3normalise(v3sub(Lp, P));
    float diff = fabsf(v3dot(N, lightDir));                                           // max(|N.L|, 0)
    const V3 diffuse = v3mul(Lc, v3mulf(Kd, diff));
    const V3 viewDir = v3normalise(v3sub(origin, P));
    const V3 I = { -viewDir.x, -viewDir.y, -viewDir.z };
    const V3 reflectDir = v3sub(I, v3muld(N, __dmul_rn(2.0, (double)v3dot(N, I))));    // reflect (Vec3.inl:120-124)
    const float vd = v3dot(viewDir, reflectDir);
    const float base = (0.0f < vd) ? vd : 0.0f;
    const float spec = (M.shininess == 1.0f) ? base : powf(base, M.shininess);        // pow(x, 1) == x exactly
    const V3 specular = v3mul(Lc, v3mulf(Ks, spec));
    V3 colour = v3add(v3add(ambient, diffuse), specular);
    // ---- shadow ray (:680-712)
    const V3 sdir0 = v3normalise(v3sub(Lp, P));
    const float sl = __fsqrt_rn(v3dot(sdir0, sdir0));
    V3 sdir = { 0.0f, 0.0f, 0.0f };
    if (sl != 0.0f) sdir = { __fdiv_rn(sdir0.x, sl), __fdiv_rn(sdir0.y, sl), __fdiv_rn(sdir0.z, sl) };   // Ray::Ray
    bool in_shadow = false;
    // blockers: the triangles of the ray's cell of the light-space grid + the short list every ray tests; brute force
    // over the scene when there is no grid or the hit point has no projection inside it
    auto blocked_by = [&](uint32_t g) {
        if (g == (uint32_t)id) return false;                                      // &triangle != p_intersected_triangle
        const ShadeMesh S = a.meshes[mesh_of(a.meshes, a.n_meshes, g)];
        const float* v = S.vertices + 9 * (size_t)(g - S.first);
        float ts;
        return intersect_generic(P, sdir, ld_v3(v), ld_v3(v + 3), ld_v3(v + 6), ts) && (double)ts > 0.0000001;
    };
    bool use_grid = false;
    uint32_t cell = 0;
    if (a.grid.enabled) {
        float pa, pb, dp;
        if (light_project(a.grid, a.light_position, P.x, P.y, P.z, pa, pb, dp)) {
            const float fa = floorf((pa - a.grid.a0) * a.grid.inv_da), fb = floorf((pb - a.grid.b0) * a.grid.inv_db);
            if (fa >= 0.0f && fb >= 0.0f && fa < (float)a.grid.G && fb < (float)a.grid.G) {
                use_grid = true;
                cell = (uint32_t)fb * a.grid.G + (uint32_t)fa;
            }
        }
    }
    if (use_grid) {
        const uint32_t e0 = a.grid.cell_off[cell], e1 = a.grid.cell_off[cell + 1];
        for (uint32_t e = e0; e < e1 && !in_shadow; ++e) in_shadow = blocked_by(a.grid.cell_tris[e]);
        const uint32_t ng = *a.grid.n_global;
        for (uint32_t e = 0; e < ng && !in_shadow; ++e) in_shadow = blocked_by(a.grid.global_tris[e]);
    } else {
        for (uint32_t g = 0; g < a.grid.total_tris && !in_shadow; ++g) in_shadow = blocked_by(g);
    }
    if (a.shadow) a.shadow[pixel] = in_shadow ? 1 : 0;
    if (in_shadow) {                                                                   // colour *= 0.25 (double)
        colour.x = (float)__dmul_rn((double)colour.x, 0.25); colour.y = (float)__dmul_rn((double)colour.y, 0.25);
        colour.z = (float)__dmul_rn((double)colour.z, 0.25);
    }
    // ---- texture (:722-774) or plain conversion (:776-790)
    if (M.tex && M.tex_w * M.tex_h && M.texcoords) {
        const V3 Pt = v3add(origin, v3mulf(dirA, t));                                  // uses the pre-Ray direction (:729)
        const float* v = M.vertices + 9 * (size_t)local;
        const V3 A = ld_v3(v), B = ld_v3(v + 3), C = ld_v3(v + 6);
        const float area = tri_area(A, B, C);
        const float u = __fdiv_rn(tri_area(C, A, Pt), area);
        const float vv = __fdiv_rn(tri_area(A, B, Pt), area);
        const float w = __fdiv_rn(tri_area(B, C, Pt), area);
        const float* tc = M.texcoords + 9 * (size_t)local;
        const V3 texel = v3add(v3add(v3mulf(ld_v3(tc), w), v3mulf(ld_v3(tc + 3), u)), v3mulf(ld_v3(tc + 6), vv));
        uint32_t ti = (uint32_t)__fmul_rn(texel.x, (float)(M.tex_w - 1)), tj = (uint32_t)__fmul_rn(texel.y, (float)(M.tex_h - 1));
        if (ti >= M.tex_w) ti = M.tex_w - 1;                                          // the reference would throw here
        if (tj >= M.tex_h) tj = M.tex_h - 1;
        const uint8_t* px = M.tex + ((size_t)tj * M.tex_w + ti) * 3;
        colour.x = __fmul_rn(colour.x, (float)px[0]); colour.y = __fmul_rn(colour.y, (float)px[1]);
        colour.z = __fmul_rn(colour.z, (float)px[2]);
        out[0] = clamp_trunc(colour.x); out[1] = clamp_trunc(colour.y); out[2] = clamp_trunc(colour.z);
    } else {
        out[0] = clamp_trunc(__dmul_rn(255.0, (double)colour.x)); out[1] = clamp_trunc(__dmul_rn(255.0, (double)colour.y));
        out[2] = clamp_trunc(__dmul_rn(255.0, (double)colour.z));
    }
}

// ------------------------------------------------------------------ de-interleave
// slots[part][local_row][col] with local_row = (block / n_parts) * 16 + r holds image row block * 16 + r of
// block = local_block * n_parts + part.  float4 copies when W % 4 == 0.
__global__ void __launch_bounds__(256) deinterleave_kernel(const float* __restrict__ slots, float* __restrict__ frame,
                                                           uint32_t W, uint32_t H, uint32_t n_parts, uint32_t slot_rows)
{
    const uint32_t row = blockIdx.y;
    const uint32_t blk = row >> 4, r = row & 15u;
    const uint32_t part = blk % n_parts, lrow = (blk / n_parts) * 16 + r;
    const float* src = slots + ((size_t)part * slot_rows + lrow) * W;
    float* dst = frame + (size_t)row * W;
    if ((W & 3u) == 0) {
        for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) * 4; c < W; c += gridDim.x * blockDim.x * 4)
            *reinterpret_cast<float4*>(dst + c) = *reinterpret_cast<const float4*>(src + c);
    } else {
        for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < W; c += gridDim.x * blockDim.x) dst[c] = src[c];
    }
}

// ------------------------------------------------------------------ repair kernel
// lbuffer.cxx:329-404.  Index arithmetic is unsigned 32-bit like the reference's
// (`col - i`, `row - i` wrap); the reference's `index > size` guard would read one
// element past the end when index == size, here that case stops the probe.
__device__ __forceinline__ float repair_probe_dir(const float* __restrict__ L, uint32_t base, int32_t step,
                                                  uint32_t depth, uint64_t size)
{
    float value = 0.0f;
    for (uint32_t i = 1; i <= depth; ++i) {
        const uint32_t index = base + (uint32_t)((int32_t)i * step);
        if ((uint64_t)index >= size) break;
        const float v = L[index];
        if (v != -1.0f) { value = v; break; }
    }
    return value;
}

__device__ __forceinline__ float repair_value(const float* __restrict__ L, uint64_t pixel, uint32_t W, uint64_t size, uint32_t depth)
{
    float out = L[pixel];
    if (out == -1.0f) {
        const uint32_t row = (uint32_t)(pixel / W), col = (uint32_t)(pixel % W);
        const uint32_t base = row * W + col;
        float sum = 0.0f; uint32_t n = 0;
        float v;
        v = repair_probe_dir(L, base, 1, depth, size);               if (v != 0.0f) { sum = __fadd_rn(sum, v); ++n; }   // +col
        v = repair_probe_dir(L, base, -(int32_t)W, depth, size);     if (v != 0.0f) { sum = __fadd_rn(sum, v); ++n; }   // -row
        v = repair_probe_dir(L, base, -1, depth, size);              if (v != 0.0f) { sum = __fadd_rn(sum, v); ++n; }   // -col
        v = repair_probe_dir(L, base, (int32_t)W, depth, size);      if (v != 0.0f) { sum = __fadd_rn(sum, v); ++n; }   // +row
        out = __fdiv_rn(sum, (float)n);                              // 0/0 -> NaN like the reference
    }
    return out;
}

__global__ void __launch_bounds__(256) repair_kernel(const float* __restrict__ L, float* __restrict__ image,
                                                     uint32_t W, uint32_t H, uint32_t depth)
{
    const uint64_t size = (uint64_t)W * H;
    const uint64_t pixel = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= size) return;
    image[pixel] = repair_value(L, pixel, W, size, depth);
}

// Sharded repair (multi-GPU): the rank repairs the pixels of ITS 16-row blocks only -- reading the assembled L frame, whose
// rows around the blocks came from the neighbouring ranks -- and stores them into the image of every rank (peer
// mappings, or one multimem.st into the NVSwitch multicast mapping).  4 pixels per thread, 128-bit stores.
__global__ void __launch_bounds__(256) repair_blocks_kernel(const float* __restrict__ L, const RepairOut out, uint32_t W, uint32_t H,
                                                            uint32_t depth, uint32_t n_parts, uint32_t part)
{
    const uint32_t by = blockIdx.y >> 4, lrow = blockIdx.y & 15u;
    const uint32_t row = (by * n_parts + part) * 16 + lrow;
    if (row >= H) return;
    const uint64_t size = (uint64_t)W * H;
    auto store1 = [&](uint64_t p, float v) {
        if (out.multicast) asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(out.image + p), "f"(v) : "memory");
        else out.image[p] = v;
        for (uint32_t i = 0; i < out.n_extra; ++i) out.extra[i][p] = v;
    };
    if (out.vec4) {
        const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (c >= W) return;
        const uint64_t p = (uint64_t)row * W + c;
        const float4 v = make_float4(repair_value(L, p, W, size, depth), repair_value(L, p + 1, W, size, depth),
                                     repair_value(L, p + 2, W, size, depth), repair_value(L, p + 3, W, size, depth));
        if (out.multicast)
            asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out.image + p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else *reinterpret_cast<float4*>(out.image + p) = v;
        for (uint32_t i = 0; i < out.n_extra; ++i) *reinterpret_cast<float4*>(out.extra[i] + p) = v;
    } else {
        for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < W; c += gridDim.x * blockDim.x)
            store1((uint64_t)row * W + c, repair_value(L, (uint64_t)row * W + c, W, size, depth));
    }
}

// ------------------------------------------------------------------ scene pre-pack
// One thread per slot of the chunk-major scan array (>= the padded triangle count).  Exact constants use
// round-to-nearest intrinsics (prepack_math.h); the filter coefficients / margins are double precision.  The level-0
// record is written twice: in natural triangle order (filt0: queue stages, bin pass, culled kernel) and into the
// chunk-major SoA array the filtered kernel bulk-copies: slot i = (k 128 + t) n_chunks + c  ->  scan[c][k][j][t].
__global__ void __launch_bounds__(128) prepack_kernel(const float* __restrict__ vertices, const float* __restrict__ normals,
                                                      uint32_t n, uint32_t n_padded, uint32_t n_chunks, uint32_t scan_k,
                                                      float ox, float oy, float oz,
                                                      const __grid_constant__ FrameGeom g, ExactTri* exact,
                                                      float4* filt0, float4* filt1, float4* scan)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks * scan_k * 128) return;
    FilterTri f; FilterTri2 f2;
    if (i < n) {
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = vertices[9 * (size_t)i + k];
        float nrm[3];
        const float origin[3] = { ox, oy, oz };
        if (normals) { nrm[0] = normals[3 * (size_t)i]; nrm[1] = normals[3 * (size_t)i + 1]; nrm[2] = normals[3 * (size_t)i + 2]; }
        ExactTri E;
        pack_triangle(v, normals ? nrm : nullptr, origin, g, E, f, f2);
        float4* eo = reinterpret_cast<float4*>(exact + i);
        eo[0] = make_float4(E.e2x, E.e2y, E.e2z, E.e1x); eo[1] = make_float4(E.e1y, E.e1z, E.tx, E.ty);
        eo[2] = make_float4(E.tz, E.qx, E.qy, E.qz);     eo[3] = make_float4(E.e2q, E.nx, E.ny, E.nz);
    } else pack_padding(f, f2);
    const float4 r0 = make_float4(f.bx, f.gx, f.ax, f.bd), r1 = make_float4(f.gd, f.ad, f.m, f.R), r2 = make_float4(f.bx2, f.gx2, f.ax2, f.R2);
    if (i < n_padded) {
        filt0[3 * (size_t)i] = r0; filt0[3 * (size_t)i + 1] = r1; filt0[3 * (size_t)i + 2] = r2;
        filt1[2 * (size_t)i] = make_float4(f2.bx2, f2.gx2, f2.ax2, f2.m2);  filt1[2 * (size_t)i + 1] = make_float4(f2.bx3, f2.gx3, f2.ax3, f2.m3);
    }
    const uint32_t c = i % n_chunks, q = i / n_chunks, k = q / 128, t = q % 128;
    float4* dst = scan + ((size_t)(c * scan_k + k) * 3) * 128 + t;
    dst[0] = r0; dst[128] = r1; dst[256] = r2;
}

// Per-column / per-row detector offsets exactly as lbuffer.cxx:756-757:
//   float v_offset = pixel_spacing[1] * (0.5 + row - H / 2.0)   (double product, narrowed to float)
__global__ void offsets_kernel(float* u_off, float* v_off, uint32_t W, uint32_t H, float sp_u, float sp_v)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W) u_off[i] = (float)__dmul_rn((double)sp_u, (0.5 + (double)i - (double)W / 2.0));
    if (i < H) v_off[i] = (float)__dmul_rn((double)sp_v, (0.5 + (double)i - (double)H / 2.0));
}

// ------------------------------------------------------------------ FFMA peak probe
// 16 independent FFMA chains per thread: the measured FP32 roofline denominator (bench.py).
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b)
{
    float r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = (float)(threadIdx.x + k);
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < 16; ++k) r[k] = fmaf(r[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += r[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace

cudaError_t launch_prepack(const float* vertices, const float* normals, uint32_t n, uint32_t n_padded, uint32_t n_chunks,
                           uint32_t scan_k, const float origin[3], const FrameGeom& g, void* exact, void* filt0, void* filt1,
                           void* scan, cudaStream_t s)
{
    const uint32_t slots = n_chunks * scan_k * 128;
    if (slots == 0) return cudaSuccess;
    prepack_kernel<<<(slots + 127) / 128, 128, 0, s>>>(vertices, normals, n, n_padded, n_chunks, scan_k, origin[0], origin[1],
                                                       origin[2], g, (ExactTri*)exact, (float4*)filt0, (float4*)filt1, (float4*)scan);
    return cudaGetLastError();
}

cudaError_t launch_offsets(float* u_off, float* v_off, uint32_t W, uint32_t H, float sp_u, float sp_v, cudaStream_t s)
{
    const uint32_t n = W > H ? W : H;
    if (n == 0) return cudaSuccess;
    offsets_kernel<<<(n + 255) / 256, 256, 0, s>>>(u_off, v_off, W, H, sp_u, sp_v);
    return cudaGetLastError();
}

cudaError_t launch_ffma_peak(float* scratch, int blocks, int iters, cudaStream_t s)
{
    ffma_peak_kernel<<<blocks, 256, 0, s>>>(scratch, iters, 0.999f, 0.001f);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ peer barrier (one process per GPU)
// Device-side barrier over the ranks of one node, on flags in peer-mapped memory: thread j publishes this rank's epoch
// in rank j's flag array and waits until rank j's epoch has arrived in this rank's.  The kernels launched before it on
// the stream have completed, so their stores into the peers' frames (plain, peer or multimem) are ordered before the
// flag by the system-scope fence; the kernels launched after it (the repair pass) start when every rank has signalled,
// i.e. when every rank's rows have landed in this rank's frame.  One CTA: a waiting barrier never occupies more than
// one CTA slot, so barriers of several frames in flight (one per stream) cannot starve the kernels they wait for.
// phases: bit 0 = signal, bit 1 = wait (3 = the whole barrier in one launch).
// ctr != nullptr: the epoch is the next value of a counter in device memory (one barrier at a time uses it: barriers on a
// set of flags are ordered by their stream), so that a captured CUDA graph of a frame can be replayed as it is.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerFlags f, uint32_t n, uint32_t rank, uint32_t epoch, uint32_t phases,
                                                          uint32_t* ctr)
{
    __shared__ uint32_t s_epoch;
    if (ctr) {
        if (threadIdx.x == 0) { s_epoch = *ctr + 1; *ctr = s_epoch; }
        __syncthreads();
        epoch = s_epoch;
    }
    const uint32_t j = threadIdx.x;
    if (j >= n || j == rank) return;
    __threadfence_system();
    if (phases & 1u) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[j] + rank), "r"(epoch) : "memory");
    if (phases & 2u) {
        uint32_t v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.p[rank] + j) : "memory");
        } while ((int32_t)(v - epoch) < 0);                                // epochs only grow (wrap-safe comparison)
        __threadfence_system();
    }
}

cudaError_t launch_peer_barrier(const PeerFlags& f, uint32_t n, uint32_t rank, uint32_t epoch, uint32_t phases, uint32_t* ctr, cudaStream_t s)
{
    if (n <= 1) return cudaSuccess;
    peer_barrier_kernel<<<1, 32, 0, s>>>(f, n, rank, epoch, phases, ctr);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_lbuffer_exact(const RenderArgs& a, int audit, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0) return cudaSuccess;
    dim3 grid((a.cam.W + 7) / 8, a.n_blocks);
    if (audit) lbuffer_exact_kernel<true><<<grid, 128, 0, s>>>(a);
    else lbuffer_exact_kernel<false><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

// shapes the filtered kernel is built for: {box width, box height, triangles per thread}
#define SRT_SCAN_SHAPES(X) X(16, 2, 4) X(16, 1, 4) X(32, 1, 4) X(8, 4, 4) X(8, 1, 4) X(32, 2, 4) X(16, 4, 4) X(8, 8, 4) X(16, 2, 2) X(16, 2, 6) X(32, 2, 6) \
    X(32, 4, 4) X(32, 8, 4) X(32, 16, 4) X(64, 8, 4) X(64, 16, 4) X(16, 16, 4)

// Dynamic shared memory of a scan CTA.  scan_max_ctas (0 = no cap) caps the CTAs resident per SM by padding the
// request: with frames pipelined on several streams, a scan that leaves a quarter of the register file free lets the
// latency-bound tail kernels of the previous frame run beside it instead of behind it.
static size_t scan_smem_bytes(size_t need, uint32_t max_ctas)
{
    if (max_ctas == 0) return need;
    const size_t per_sm = 228 * 1024, reserved = 1024;                    // sm_100: 228 KB per SM, 1 KB reserved per CTA
    const size_t pad = per_sm / (max_ctas + 1) - reserved + 256;          // max_ctas + 1 CTAs of this size do not fit
    return std::min<size_t>(std::max(need, pad), 48 * 1024);
}

template <int PX, int PY, int K>
static cudaError_t launch_filtered_t(const RenderArgs& a, cudaStream_t s)
{
    using G = ScanGeom<PX, PY, K>;
    static_assert(G::kSmem <= 48 * 1024, "the chunk and the queues must fit the default dynamic shared-memory limit");
    lbuffer_filtered_kernel<PX, PY, K><<<a.scan_ctas, 128, scan_smem_bytes(G::kSmem, a.scan_max_ctas), s>>>(a);
    return cudaGetLastError();
}

template <int PX, int PY>
static cudaError_t launch_candidates_t(const RenderArgs& a, cudaStream_t s)
{
    const int sms = a.sm_count ? (int)a.sm_count : 148;
    // grid-stride over the box list / the pixel list; when the host knows the list sizes of this launch geometry
    // (verified on its first render) the grids are sized to them: a share of a multi-GPU frame has short lists
    const unsigned full = (unsigned)sms * 32;
    const unsigned g1 = a.cand_hint ? std::min(full, (a.cand_hint + 127) / 128 + 1) : full;
    const unsigned g2 = a.pix_hint ? std::min(full, (a.pix_hint + 127) / 128 + 1) : full;
    cudaError_t e;
    if constexpr (ScanGeom<PX, PY, 4>::kBig) {
        if (!a.big || !a.big_count) return cudaErrorInvalidValue;
        const unsigned g0 = a.big_hint ? std::min(full, (a.big_hint + 127) / 128 + 1) : full;
        refine_kernel<PX, PY><<<g0, 128, 0, s>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        candidates_kernel<16, 4><<<g1, 128, 0, s>>>(a);
    } else {
        candidates_kernel<PX, PY><<<g1, 128, 0, s>>>(a);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    exact_tests_kernel<<<g2, 128, 0, s>>>(a);
    return cudaGetLastError();
}

bool filtered_shape_supported(uint32_t px, uint32_t py, uint32_t k)
{
#define X(PX, PY, K) if (px == PX && py == PY && k == K) return true;
    SRT_SCAN_SHAPES(X)
#undef X
    return false;
}

template <int PX, int PY, int K>
static int filtered_occupancy_t()
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lbuffer_filtered_kernel<PX, PY, K>, 128, ScanGeom<PX, PY, K>::kSmem) != cudaSuccess) n = 0;
    return n;
}

int filtered_resident_ctas(uint32_t px, uint32_t py, uint32_t k, uint32_t max_ctas)
{
#define X(PX, PY, K) if (px == PX && py == PY && k == K) { const int n = filtered_occupancy_t<PX, PY, K>(); return (max_ctas && n > (int)max_ctas) ? (int)max_ctas : n; }
    SRT_SCAN_SHAPES(X)
#undef X
    return 0;
}

cudaError_t launch_filtered_candidates(const RenderArgs& a, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0 || a.scan_chunks == 0) return cudaSuccess;
#define X(PX, PY, K) if (a.box_px == PX && a.box_py == PY && a.scan_k == K) return launch_candidates_t<PX, PY>(a, s);
    SRT_SCAN_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_lbuffer_filtered(const RenderArgs& a, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0 || a.scan_chunks == 0) return cudaSuccess;
    if (!a.hit_head || !a.hit_next || a.scan_cols == 0 || a.scan_parts_x == 0 || a.scan_ctas == 0 || !a.cand || !a.cand_count || !a.pix ||
        a.n_meshes > 32)
        return cudaErrorInvalidValue;
#define X(PX, PY, K) if (a.box_px == PX && a.box_py == PY && a.scan_k == K) return launch_filtered_t<PX, PY, K>(a, s);
    SRT_SCAN_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_lbuffer_bins(const RenderArgs& a, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0) return cudaSuccess;
    bin_kernel<<<dim3(a.bins_x, a.bins_y, (unsigned)a.n_meshes), kBinThreads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_lbuffer_culled(const RenderArgs& a, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0) return cudaSuccess;
    lbuffer_culled_kernel<<<dim3((a.cam.W + 63) / 64, a.n_blocks), 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_resolve_hits(const RenderArgs& a, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0) return cudaSuccess;
    dim3 grid((a.cam.W + 255) / 256, a.n_blocks * 16);
    resolve_hits_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scene_bbox(const ShadeMesh* meshes, int n_meshes, uint32_t total_tris, uint32_t* out6, cudaStream_t s)
{
    if (total_tris == 0) return cudaSuccess;
    scene_bbox_kernel<<<(unsigned)std::min<uint64_t>(1184, ((uint64_t)total_tris + 255) / 256), 256, 0, s>>>(meshes, n_meshes, total_tris, out6);
    return cudaGetLastError();
}

float ordered_key_to_float(uint32_t k)
{
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &b, sizeof f);
    return f;
}

cudaError_t launch_shadow_bin(const ShadeArgs& a, int fill, cudaStream_t s)
{
    if (a.grid.total_tris == 0) return cudaSuccess;
    const unsigned blocks = (a.grid.total_tris + 255) / 256;
    if (fill) shadow_bin_kernel<true><<<blocks, 256, 0, s>>>(a);
    else shadow_bin_kernel<false><<<blocks, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_exclusive_scan(uint32_t* data, uint32_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(data, n);
    return cudaGetLastError();
}

cudaError_t launch_shade(const ShadeArgs& a, cudaStream_t s)
{
    if (a.row_end <= a.row_begin || a.cam.W == 0) return cudaSuccess;
    dim3 grid((a.cam.W + 7) / 8, (a.row_end - a.row_begin + 15) / 16);
    shade_kernel<<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_deinterleave(const float* slots, float* frame, uint32_t W, uint32_t H, uint32_t n_parts,
                                uint32_t slot_rows, cudaStream_t s)
{
    if (W == 0 || H == 0) return cudaSuccess;
    dim3 grid((W / 4 + 255) / 256 ? (W / 4 + 255) / 256 : 1, H);
    deinterleave_kernel<<<grid, 256, 0, s>>>(slots, frame, W, H, n_parts, slot_rows);
    return cudaGetLastError();
}

cudaError_t launch_repair_blocks(const float* L, const RepairOut& out, uint32_t W, uint32_t H, uint32_t depth, uint32_t n_parts,
                                 uint32_t part, cudaStream_t s)
{
    if (W == 0 || H == 0 || n_parts == 0) return cudaSuccess;
    const uint32_t blocks = (H + 15) / 16;
    const uint32_t mine = blocks > part ? (blocks - part + n_parts - 1) / n_parts : 0;
    if (mine == 0) return cudaSuccess;
    dim3 grid(out.vec4 ? (W / 4 + 255) / 256 : std::min(8u, (W + 255) / 256), mine * 16);
    repair_blocks_kernel<<<grid, 256, 0, s>>>(L, out, W, H, depth, n_parts, part);
    return cudaGetLastError();
}

cudaError_t launch_repair(const float* L, float* image, uint32_t W, uint32_t H, uint32_t depth, cudaStream_t s)
{
    const uint64_t size = (uint64_t)W * H;
    if (size == 0) return cudaSuccess;
    repair_kernel<<<(unsigned)((size + 255) / 256), 256, 0, s>>>(L, image, W, H, depth);
    return cudaGetLastError();
}

} // namespace srt
