// prepack_math.h -- per-triangle pre-pack arithmetic shared by the host (prepack.cpp, used by the
// diagnostic srt_debug_prepack) and the device (prepack kernel in kernels.cu, used by every render).
//
// Two kinds of arithmetic live here:
//  * EXACT float operations that must equal the reference's separately rounded binary32 ops
//    (edge vectors, tvec, qvec, edge2.qvec, the triangle normal: src/Ray.cxx:86-122,
//    include/Triangle.inl:170-178, include/Vec3.inl:313-329, 461-476).  They go through the SRT_F*
//    macros: round-to-nearest intrinsics on the device (never contracted), plain operators on the
//    host (the host translation units are compiled with -ffp-contract=off).
//  * double-precision bound arithmetic for the conservative filter (coefficients and margins).
//    Its results only have to be valid upper bounds, so contraction / op order do not matter.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define SRT_HD __host__ __device__ __forceinline__
#else
#define SRT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SRT_FADD(a, b) __fadd_rn((a), (b))
#define SRT_FSUB(a, b) __fsub_rn((a), (b))
#define SRT_FMUL(a, b) __fmul_rn((a), (b))
#define SRT_FDIV(a, b) __fdiv_rn((a), (b))
#define SRT_FSQRT(a) __fsqrt_rn((a))
#else
#define SRT_FADD(a, b) ((a) + (b))
#define SRT_FSUB(a, b) ((a) - (b))
#define SRT_FMUL(a, b) ((a) * (b))
#define SRT_FDIV(a, b) ((a) / (b))
#define SRT_FSQRT(a) std::sqrt((a))
#endif

namespace srt {

// 64 B: everything the exact Moller-Trumbore sequence needs (4 x float4).
struct ExactTri {
    float e2x, e2y, e2z, e1x;
    float e1y, e1z, tx, ty;
    float tz, qx, qy, qz;
    float e2q, nx, ny, nz;       // e2q = edge2 . qvec (Ray.cxx:122 numerator), n = Triangle::getNormal()
};

// 32 B: level-0 conservative reject.  Y and D are affine in the centred pixel coordinates (x, y):
// Y = bx*x + gx*y + ax (Y = 2 X - det for one barycentric numerator X), D = bd*x + gd*y + ad
// (D = LAMBDA * det, see pack_triangle).  A pair can only be a hit when |Y| <= |D| + m
// (<=> 0 <= X / det <= 1 up to rounding).  The record is scaled by a per-triangle power of two so
// that the render kernel can step Y and D along a pixel run in binary16 (HFMA2, two pixels per
// instruction) from a binary32 value at the run's first pixel; h holds the binary16 x-slope of Y and the gate's
// binary16 margin (the kernel steps only Y and bounds |D| over the run by |D(x0)| + 7 |bd|, folded into that margin).
struct FilterTri {
    float bx, gx, ax, bd;
    float gd, ad, m;
    uint32_t h;                  // binary16(bx) in the low half, the gate's binary16 margin mh in the high half
};

// 32 B: level-1 reject, the two remaining strips (slow path only).
struct FilterTri2 {
    float bx2, gx2, ax2, m2;
    float bx3, gx3, ax3, m3;
};

// Image-plane geometry the filter bounds depend on (per camera).
struct FrameGeom {
    double origin[3], F[3];      // F = detector - origin (real arithmetic on the float inputs)
    double up[3], right[3];
    double sp_u, sp_v;           // pixel spacing columns / rows
    double xmax, ymax;           // |x| <= xmax, |y| <= ymax over the frame (centred pixel coordinates)
    double delta[3];             // abs. error bound of the reference's float ray direction before normalising
    double G[3];                 // component-wise bound of that direction
    double Kmax;                 // bound of |d1| * |dirA|
    int    filterable;           // camera passes the sanity preconditions of the proof
};

namespace pm {

constexpr double EPS = 5.9604644775390625e-08;   // 2^-24, unit roundoff of binary32
constexpr double SAFETY = 2.0;                   // covers second-order terms of the first-order bounds
constexpr double TINY = 1e-30;                   // absolute slack for gradual underflow
constexpr double BIG = 1073741824.0;             // 2^30: magnitude precondition
constexpr double LIM = 1e30;
constexpr double LAMBDA = 1.00390625;            // 1 + 2^-8: |D| is widened by this factor; pays for the binary16 relative error
constexpr int    RUN_MAX = 16;                   // longest pixel run a thread of the filtered kernel steps in binary16
constexpr double H_SLOPE = (2.0 * (RUN_MAX - 1) + 0.1) / 2048.0;   // binary16 error per unit x-slope over a run of <= RUN_MAX pixels (see pack_triangle)
constexpr double H_ABS = 9.5367431640625e-07;    // 2^-20: binary16 subnormal roundings

struct f3 { float x, y, z; };
struct d3 { double x, y, z; };

SRT_HD f3 sub(f3 a, f3 b) { return { SRT_FSUB(a.x, b.x), SRT_FSUB(a.y, b.y), SRT_FSUB(a.z, b.z) }; }
// Vec3::crossProduct, Vec3.inl:321-329
SRT_HD f3 cross(f3 a, f3 b)
{
    return { SRT_FSUB(SRT_FMUL(a.y, b.z), SRT_FMUL(a.z, b.y)), SRT_FSUB(SRT_FMUL(a.z, b.x), SRT_FMUL(a.x, b.z)),
             SRT_FSUB(SRT_FMUL(a.x, b.y), SRT_FMUL(a.y, b.x)) };
}
// Vec3::dotProduct, Vec3.inl:313-317 (left to right)
SRT_HD float dot(f3 a, f3 b) { return SRT_FADD(SRT_FADD(SRT_FMUL(a.x, b.x), SRT_FMUL(a.y, b.y)), SRT_FMUL(a.z, b.z)); }

SRT_HD d3 D(f3 a) { return { (double)a.x, (double)a.y, (double)a.z }; }
SRT_HD d3 dcross(d3 a, d3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
SRT_HD d3 dabs(d3 a) { return { std::fabs(a.x), std::fabs(a.y), std::fabs(a.z) }; }
SRT_HD double ddot(d3 a, const double b[3]) { return a.x * b[0] + a.y * b[1] + a.z * b[2]; }
SRT_HD double absdot(d3 a, const double b[3]) { return std::fabs(a.x) * b[0] + std::fabs(a.y) * b[1] + std::fabs(a.z) * b[2]; }
SRT_HD bool finite3(f3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
SRT_HD double maxabs(f3 a) { return std::fmax(std::fabs((double)a.x), std::fmax(std::fabs((double)a.y), std::fabs((double)a.z))); }

// float -> IEEE binary16 bit pattern, round to nearest even (same result as cvt.rn.f16.f32)
SRT_HD uint32_t f16_bits(float f)
{
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(f);
#else
    static_assert(sizeof(u) == sizeof(f), "binary32 expected");
    __builtin_memcpy(&u, &f, sizeof(u));
#endif
    const uint32_t sign = (u >> 16) & 0x8000u, a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return sign | 0x7e00u;                 // NaN
    if (a >= 0x477ff000u) return sign | 0x7c00u;                // >= 65520 (or inf) -> inf
    if (a < 0x38800000u) {                                      // below 2^-14: binary16 subnormal, spacing 2^-24
#if defined(__CUDA_ARCH__)
        const float af = __uint_as_float(a);
#else
        float af; __builtin_memcpy(&af, &a, sizeof(af));
#endif
        return sign | (uint32_t)std::rint((double)af * 16777216.0);     // exact product, ties to even; 1024 == smallest normal
    }
    uint32_t h = (((a >> 23) - 112u) << 10) | ((a & 0x7fffffu) >> 13);
    const uint32_t rem = a & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;     // a carry moves into the exponent, as it should
    return sign | h;
}

SRT_HD float round_up(double v)
{
    float f = (float)v;
    if ((double)f < v) f = std::nextafter(f, INFINITY);
    return f;
}

// Affine functional N . d1(x, y) with d1 = F + up*sp_v*y + right*sp_u*x, plus its error budget.
struct Affine {
    double a, b, g;      // alpha (constant), beta (x), gamma (y)
    double T;            // bound on |float-evaluated affine - kappa^-1 * reference float numerator|
    double absmax;       // bound on |value| over the frame
    double grad2;        // b^2 + g^2: (pixels^-1 scale of the strip)^2
};

SRT_HD Affine make_affine(d3 N, double r_float_path, const FrameGeom& g)
{
    Affine A;
    A.a = ddot(N, g.F);
    A.b = g.sp_u * ddot(N, g.right);
    A.g = g.sp_v * ddot(N, g.up);
    A.absmax = std::fabs(A.a) + std::fabs(A.b) * g.xmax + std::fabs(A.g) * g.ymax;
    // (1) float evaluation fma(b, x, fma(g, y, a)) with coefficients rounded to float
    const double e_aff = 3.0 * EPS * A.absmax;
    // (2) the reference's un-normalised direction differs from d1(x, y) by <= delta (absolute),
    //     and the two normalisations add a component-wise relative error <= 8 eps
    const double Gabs[3] = { g.delta[0] + 8.0 * EPS * g.G[0], g.delta[1] + 8.0 * EPS * g.G[1], g.delta[2] + 8.0 * EPS * g.G[2] };
    const double e_dir = absdot(N, Gabs);
    // (3) rounding inside the reference's float cross/dot sequence, scaled back by 1/kappa <= Kmax
    A.T = e_aff + e_dir + r_float_path * g.Kmax;
    A.grad2 = A.b * A.b + A.g * A.g;
    return A;
}

// margin of the strip test |Y| <= |D| + m (Y = 2 X - det): a hit has |2 x_e - det_e| <= |det_e| (1 + 16 eps)
SRT_HD double strip_margin(const Affine& Y, const Affine& Dn)
{
    const double m = Y.T + (1.0 + 16.0 * EPS) * Dn.T + 16.0 * EPS * Dn.absmax + 4.0 * EPS * (Dn.absmax + Y.absmax);
    return SAFETY * m + TINY;
}

} // namespace pm

// Triangle::computeNormal (include/Triangle.inl:170-178), exact op order.
SRT_HD void compute_normal(const float* v, float n[3])
{
    using namespace pm;
    const f3 p1 = { v[0], v[1], v[2] }, p2 = { v[3], v[4], v[5] }, p3 = { v[6], v[7], v[8] };
    const f3 c = cross(sub(p2, p1), sub(p3, p1));
    const float len = SRT_FSQRT(dot(c, c));                                           // Vec3::getLength
    n[0] = SRT_FDIV(c.x, len); n[1] = SRT_FDIV(c.y, len); n[2] = SRT_FDIV(c.z, len);   // three divides
}

// Pack one triangle. Returns false when the triangle (or the camera) fails the filter's
// preconditions: the filter entries are then set so that the triangle ALWAYS survives.
SRT_HD bool pack_triangle(const float* v, const float* normal_or_null, const float origin[3], const FrameGeom& g,
                          ExactTri& E, FilterTri& f, FilterTri2& f2)
{
    using namespace pm;
    const f3 o = { origin[0], origin[1], origin[2] };
    const f3 p1 = { v[0], v[1], v[2] }, p2 = { v[3], v[4], v[5] }, p3 = { v[6], v[7], v[8] };
    const f3 e1 = sub(p2, p1);                 // Ray.cxx:86
    const f3 e2 = sub(p3, p1);                 // Ray.cxx:87
    const f3 tv = sub(o, p1);                  // Ray.cxx:102
    const f3 qv = cross(tv, e1);               // Ray.cxx:112
    const float e2q = dot(e2, qv);             // Ray.cxx:122 (numerator of t)
    float nrm[3];
    if (normal_or_null) { nrm[0] = normal_or_null[0]; nrm[1] = normal_or_null[1]; nrm[2] = normal_or_null[2]; }
    else compute_normal(v, nrm);
    E.e2x = e2.x; E.e2y = e2.y; E.e2z = e2.z; E.e1x = e1.x;
    E.e1y = e1.y; E.e1z = e1.z; E.tx = tv.x; E.ty = tv.y;
    E.tz = tv.z; E.qx = qv.x; E.qy = qv.y; E.qz = qv.z;
    E.e2q = e2q; E.nx = nrm[0]; E.ny = nrm[1]; E.nz = nrm[2];

    bool ok = g.filterable && finite3(e1) && finite3(e2) && finite3(tv) && finite3(qv) &&
              maxabs(e1) < BIG && maxabs(e2) < BIG && maxabs(tv) < BIG && maxabs(qv) < BIG * BIG;
    if (ok) {
        const d3 E1 = D(e1), E2 = D(e2), TV = D(tv), QV = D(qv);
        // s = tvec . (d x e2) = d . (e2 x tvec);  det = e1 . (d x e2) = d . (e2 x e1)
        const d3 Ns = dcross(E2, TV), Nd = dcross(E2, E1);
        const d3 a1 = dabs(E1), a2 = dabs(E2), at = dabs(TV), aq = dabs(QV);
        const double A = at.x * (a2.y + a2.z) + at.y * (a2.x + a2.z) + at.z * (a2.x + a2.y);
        const double B = a1.x * (a2.y + a2.z) + a1.y * (a2.x + a2.z) + a1.z * (a2.x + a2.y);
        const double Cq = aq.x + aq.y + aq.z;
        const double rS = 6.0 * EPS * A + TINY, rD = 6.0 * EPS * B + TINY, rV = 4.0 * EPS * Cq + TINY;
        // Each barycentric numerator X in {s, q, w = det - s - q} must satisfy 0 <= X / det <= 1 for a hit, i.e.
        // |2 X - det| <= |det| whatever the sign of det.  Y = 2 X - det is again of the form N . d, so every level of
        // the filter is the same two-sided strip test |Y| <= |D| + m, and the three strips together are exactly
        // the (conservative) footprint of the triangle.
        // The strips are compared with D = LAMBDA * det, LAMBDA = 1 + 2^-8: |det| <= |LAMBDA det| keeps every bound below
        // valid as it stands, and the slack (LAMBDA - 1)|det| pays for the RELATIVE error of the binary16 evaluation.
        const d3 NdT = { LAMBDA * Nd.x, LAMBDA * Nd.y, LAMBDA * Nd.z };
        const Affine Dn = make_affine(NdT, LAMBDA * rD, g);
        const d3 Nys = { 2.0 * Ns.x - Nd.x, 2.0 * Ns.y - Nd.y, 2.0 * Ns.z - Nd.z };
        const d3 Nyv = { 2.0 * QV.x - Nd.x, 2.0 * QV.y - Nd.y, 2.0 * QV.z - Nd.z };
        const d3 Nw = { Nd.x - Ns.x - QV.x, Nd.y - Ns.y - QV.y, Nd.z - Ns.z - QV.z };
        const d3 Nyw = { 2.0 * Nw.x - Nd.x, 2.0 * Nw.y - Nd.y, 2.0 * Nw.z - Nd.z };
        const Affine S = make_affine(Nys, 2.0 * rS + rD, g);
        const Affine Vn = make_affine(Nyv, 2.0 * rV + rD, g);
        const Affine Wn = make_affine(Nyw, 2.0 * (rS + rD + rV) + rD, g);
        // narrowest strip first (largest gradient => fewest surviving pixels): sort 3 by grad2, descending
        Affine s0 = S, s1 = Vn, s2 = Wn, t;
        if (s1.grad2 > s0.grad2) { t = s0; s0 = s1; s1 = t; }
        if (s2.grad2 > s1.grad2) { t = s1; s1 = s2; s2 = t; }
        if (s1.grad2 > s0.grad2) { t = s0; s0 = s1; s1 = t; }
        const double m0 = strip_margin(s0, Dn), m1 = strip_margin(s1, Dn), m2 = strip_margin(s2, Dn);
        // --- binary16 stepping of level 0 (render kernel, even pixels-per-thread).  With Yf(k) the binary32 value
        // fma(bx, x0 + k, fma(gx, y, ax)) the bounds above are about, the kernel evaluates, for pixel k < RUN_MAX of a run,
        //     Y16(k) = fl16(fl16(bx) * k + fl16(Yf(0)))            (one HFMA2; k is exact in binary16)
        // |Y16(k) - Yf(k)| <= k |fl16(bx) - bx| + |fl16(Yf(0)) - Yf(0)| + 2 eps absmax + 2^-11 |Y16(k)| + subnormal terms
        //                  <= rho |Yf(k)| + hY,   rho = 1.001 * 2^-10,  hY = ((2 k_max + 0.1) / 2048) |bx| + 2.1 eps absmax + 2^-21
        // (|Yf(0)| <= |Yf(k)| + k_max |bx| + ..., k_max = RUN_MAX - 1 = 15), and the same for D.  The float-level statement is
        //     hit  =>  |Yf| <= (1 + 16 eps) / LAMBDA * |Df| + m0,
        // hence hit => |Y16| <= (1 + rho)(1 + 16 eps) / (LAMBDA (1 - rho)) (|D16| + hD) + (1 + rho) m0 + hY
        //                    <= |D16| + hD + 1.002 m0 + hY                    [the factor is < 1 for LAMBDA = 1 + 2^-8]
        // The record is scaled by a power of two `sc` (exact, the tests are scale invariant) so that a hit implies
        // |D16| + margin <= 2^13 and 15 |bx| <= 15 * 2^11: then no binary16 quantity of a HIT overflows, and an overflowing
        // Y16 (= +-inf, never NaN: fl16(bx) * k stays finite) is a correct reject because D16 stays finite.
        const double hY = H_SLOPE * std::fabs(s0.b) + 2.1 * EPS * s0.absmax;
        const double hD = H_SLOPE * std::fabs(Dn.b) + 2.1 * EPS * Dn.absmax;
        const double M0 = 1.002 * m0 + hY + hD;
        double sc = 1.0;
        {
            const double U = Dn.absmax + M0, B = std::fabs(s0.b);
            double lim = 0.0;
            if (U > 0.0) lim = 8000.0 / U;
            if (B > 0.0) lim = lim > 0.0 ? std::fmin(lim, 2048.0 / B) : 2048.0 / B;
            if (lim > 0.0 && std::isfinite(lim)) {
                int e = (int)std::floor(std::log2(lim));
                e = e < -100 ? -100 : (e > 100 ? 100 : e);
                sc = std::ldexp(1.0, e);
                if (sc > lim) sc *= 0.5;                        // log2 rounding at an exact power of two
            }
        }
        const double M16 = (sc * M0 + H_ABS + TINY) * 1.001953125;
        ok = std::isfinite(m0) && std::isfinite(m1) && std::isfinite(m2) && sc * m0 < LIM && sc * m1 < LIM && sc * m2 < LIM &&
             sc * S.absmax < LIM && sc * Dn.absmax < LIM && sc * Vn.absmax < LIM && sc * Wn.absmax < LIM &&
             sc * (Dn.absmax + M0) <= 8192.0 && sc * std::fabs(s0.b) <= 2050.0;
        if (ok) {
            f.bx = (float)(sc * s0.b); f.gx = (float)(sc * s0.g); f.ax = (float)(sc * s0.a);
            f.bd = (float)(sc * Dn.b); f.gd = (float)(sc * Dn.g); f.ad = (float)(sc * Dn.a);
            f.m = round_up(M16);
            // binary16 margin of the kernel's gate  min_k |Y16(k)| <= tau16 = fl16(|fl16(D(x0))| + mh):  |D| over a run of
            // <= RUN_MAX pixels exceeds |D(x0)| by at most (RUN_MAX - 1) |bd| (+ 2 eps absmax, part of M0), fl16(D(x0)) is short by at most
            // 2^-11 |D| (LAMBDA pays for it, see above) and the final sum is rounded once more (<= 2^-11 (|D| + mh));
            // (1 + 2^-10) before the round-to-nearest conversion keeps the stored value an upper bound.
            const double mh = (M16 + (RUN_MAX - 1) * std::fabs((double)f.bd) * 1.001 + sc * Dn.absmax / 512.0 + H_ABS) * 1.0009765625;
            f.h = f16_bits(f.bx) | (f16_bits((float)mh) << 16);
            f2.bx2 = (float)(sc * s1.b); f2.gx2 = (float)(sc * s1.g); f2.ax2 = (float)(sc * s1.a); f2.m2 = round_up(sc * m1 + TINY);
            f2.bx3 = (float)(sc * s2.b); f2.gx3 = (float)(sc * s2.g); f2.ax3 = (float)(sc * s2.a); f2.m3 = round_up(sc * m2 + TINY);
        }
    }
    if (!ok) {
        // always survives: X == D == 0 -> |D| - |X| = 0 >= -m
        f = FilterTri{ 0, 0, 0, 0, 0, 0, 0, 0 };
        f2 = FilterTri2{ 0, 0, 0, 0, 0, 0, 0, 0 };
    }
    return ok;
}

// Padding entries never survive: X = 3e38 (binary16: +inf), D = 0 -> |D| - |X| < -m for every tile margin (< 1e30).
SRT_HD void pack_padding(FilterTri& f, FilterTri2& f2)
{
    f = FilterTri{ 0, 0, 3.0e38f, 0, 0, 0, 0, 0 };
    f2 = FilterTri2{ 0, 0, 3.0e38f, 0, 0, 0, 3.0e38f, 0 };
}

} // namespace srt
