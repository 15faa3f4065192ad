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

// 48 B: level-0 conservative reject.  Y and D are affine in the centred pixel coordinates (x, y):
// Y = bx*x + gx*y + ax (Y = 2 X - det for one barycentric numerator X), D = bd*x + gd*y + ad (D = det).
// A pair can only be a hit when |Y| <= |D| + m (<=> 0 <= X / det <= 1 up to rounding).  R is the same test for
// a whole BOX of pixel centres of the filtered kernel (half extents FrameGeom::box_hx / box_hy): no pixel of a box
// with centre c can hit when |Y(c)| - |D(c)| > R,  R = (hx (|bx| + |bd|) + hy (|gx| + |gd|) + 2 m)(1 + 2^-18).
// (bx2, gx2, ax2, R2) is the box test of a SECOND strip (a copy of FilterTri2's first strip): one strip is an infinite
// band across the image, two strips intersect in a parallelogram around the triangle, so boxes that pass both are rare.
struct FilterTri {
    float bx, gx, ax, bd;
    float gd, ad, m, R;
    float bx2, gx2, ax2, R2;
};

// 32 B: level-1 reject, the two remaining strips (queue stages of the filtered kernel, bin pass, culled kernel).
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
    double box_hx, box_hy;       // half extents (pixels) of the box one thread of the filtered kernel tests per triangle
    int    filterable;           // camera passes the sanity preconditions of the proof
};

namespace pm {

constexpr double EPS = 5.9604644775390625e-08;   // 2^-24, unit roundoff of binary32
constexpr double SAFETY = 2.0;                   // covers second-order terms of the first-order bounds
constexpr double TINY = 1e-30;                   // absolute slack for gradual underflow
constexpr double BIG = 1073741824.0;             // 2^30: magnitude precondition
constexpr double LIM = 1e30;
constexpr double BOX_INFLATE = 1.000003814697265625;   // 1 + 2^-18: roundings of the box test's own float arithmetic

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
        const Affine Dn = make_affine(Nd, rD, g);
        const d3 Nys = { 2.0 * Ns.x - Nd.x, 2.0 * Ns.y - Nd.y, 2.0 * Ns.z - Nd.z };
        const d3 Nyv = { 2.0 * QV.x - Nd.x, 2.0 * QV.y - Nd.y, 2.0 * QV.z - Nd.z };
        const d3 Nw = { Nd.x - Ns.x - QV.x, Nd.y - Ns.y - QV.y, Nd.z - Ns.z - QV.z };
        const d3 Nyw = { 2.0 * Nw.x - Nd.x, 2.0 * Nw.y - Nd.y, 2.0 * Nw.z - Nd.z };
        Affine st[3] = { make_affine(Nys, 2.0 * rS + rD, g), make_affine(Nyv, 2.0 * rV + rD, g),
                         make_affine(Nyw, 2.0 * (rS + rD + rV) + rD, g) };
        const float bd = (float)Dn.b, gd = (float)Dn.g, ad = (float)Dn.a;
        float mf[3]; double R[3], score[3];
        ok = std::isfinite(Dn.absmax) && Dn.absmax < LIM;
        for (int i = 0; i < 3; ++i) {
            const double m = strip_margin(st[i], Dn);
            ok = ok && std::isfinite(m) && m < LIM && st[i].absmax < LIM;
            mf[i] = round_up(m);
            // Box test of the filtered kernel (filter proof item 7): for every pixel centre p of a box with
            // centre c and half extents (hx, hy), |Y(p)| >= |Y(c)| - r_Y and |D(p)| <= |D(c)| + r_D with r = |beta| hx +
            // |gamma| hy (the float coefficients the kernel evaluates), so no pixel of the box can hit when
            // |Y(c)| - |D(c)| > r_Y + r_D + 2 m: one m for the pixel, one for the float evaluation at the centre.
            R[i] = (g.box_hx * (std::fabs((double)(float)st[i].b) + std::fabs((double)bd)) +
                    g.box_hy * (std::fabs((double)(float)st[i].g) + std::fabs((double)gd)) + 2.0 * (double)mf[i]) * BOX_INFLATE;
            // half width (pixels) of the set of box centres this strip lets through, |D| taken at the image centre
            score[i] = st[i].grad2 > 0.0 ? (std::fabs(Dn.a) + R[i]) / std::sqrt(st[i].grad2) : INFINITY;
        }
        // The filtered kernel tests two strips per box: the pair whose parallelogram of surviving box centres is the
        // smallest (area = 4 h_i h_j / |n_i x n_j| with unit normals n; a heuristic that only affects speed).  The
        // narrower of the two is level 0.
        int i0 = 0, i1 = 1; double best = INFINITY;
        for (int i = 0; i < 3; ++i)
            for (int j = i + 1; j < 3; ++j) {
                const double cr = std::fabs(st[i].b * st[j].g - st[i].g * st[j].b);        // |grad_i x grad_j|
                const double area = cr > 0.0 ? score[i] * score[j] * std::sqrt(st[i].grad2 * st[j].grad2) / cr : INFINITY;
                if (area < best || (i == 0 && j == 1)) { best = area; i0 = i; i1 = j; }
            }
        if (score[i1] < score[i0]) { const int t = i0; i0 = i1; i1 = t; }
        const int i2 = 3 - i0 - i1;
        ok = ok && std::isfinite(R[i0]) && R[i0] < LIM && std::isfinite(R[i1]) && R[i1] < LIM;
        if (ok) {
            f.bx = (float)st[i0].b; f.gx = (float)st[i0].g; f.ax = (float)st[i0].a;
            f.bd = bd; f.gd = gd; f.ad = ad;
            f.m = mf[i0]; f.R = round_up(R[i0]);
            f.bx2 = (float)st[i1].b; f.gx2 = (float)st[i1].g; f.ax2 = (float)st[i1].a; f.R2 = round_up(R[i1]);
            f2.bx2 = f.bx2; f2.gx2 = f.gx2; f2.ax2 = f.ax2; f2.m2 = mf[i1];
            f2.bx3 = (float)st[i2].b; f2.gx3 = (float)st[i2].g; f2.ax3 = (float)st[i2].a; f2.m3 = mf[i2];
        }
    }
    if (!ok) {
        // always survives: Y == D == 0 -> |Y| - |D| = 0 is never > m or > R
        f = FilterTri{ 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        f2 = FilterTri2{ 0, 0, 0, 0, 0, 0, 0, 0 };
    }
    return ok;
}

// Padding entries never survive: Y = 3e38, D = 0 -> |Y| - |D| > R = 0.
SRT_HD void pack_padding(FilterTri& f, FilterTri2& f2)
{
    f = FilterTri{ 0, 0, 3.0e38f, 0, 0, 0, 0, 0, 0, 0, 3.0e38f, 0 };
    f2 = FilterTri2{ 0, 0, 3.0e38f, 0, 0, 0, 3.0e38f, 0 };
}

} // namespace srt
