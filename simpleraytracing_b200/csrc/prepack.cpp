// prepack.cpp -- see prepack.h.  Compile with -ffp-contract=off (no FMA contraction):
// the float results below must equal the reference's (src/Ray.cxx:86-122,
// include/Triangle.inl:170-178, include/Vec3.inl:313-329, 461-476) bit for bit.
#include "prepack.h"

#include <algorithm>
#include <cmath>
#include <limits>

#if defined(__FMA__) && !defined(SRT_ALLOW_FMA_HOST)
#error "build prepack.cpp without -mfma/-march=native (contraction would change the bits)"
#endif

namespace srt {

namespace {

struct f3 { float x, y, z; };
inline f3 sub(f3 a, f3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
// Vec3::crossProduct, Vec3.inl:321-329
inline f3 cross(f3 a, f3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
// Vec3::dotProduct, Vec3.inl:313-317 (left to right)
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct d3 { double x, y, z; };
inline d3 D(f3 a) { return { (double)a.x, (double)a.y, (double)a.z }; }
inline d3 dcross(d3 a, d3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double ddot(d3 a, const double b[3]) { return a.x * b[0] + a.y * b[1] + a.z * b[2]; }
inline double absdot(d3 a, const double b[3]) { return std::fabs(a.x) * b[0] + std::fabs(a.y) * b[1] + std::fabs(a.z) * b[2]; }

constexpr double EPS = 5.9604644775390625e-08;   // 2^-24, unit roundoff of binary32
constexpr double SAFETY = 2.0;                   // covers second-order terms of the first-order bounds
constexpr double TINY = 1e-30;                   // absolute slack for gradual underflow
constexpr double BIG = 1073741824.0;             // 2^30: magnitude precondition

inline float round_up(double v)
{
    float f = (float)v;
    if ((double)f < v) f = std::nextafter(f, std::numeric_limits<float>::infinity());
    return f;
}

// Affine functional N . d1(x, y) with d1 = F + up*sp_v*y + right*sp_u*x, plus its error budget.
struct Affine {
    double a, b, g;      // alpha (constant), beta (x), gamma (y)
    double T;            // bound on |float-evaluated affine - kappa^-1 * reference float numerator|
    double absmax;       // bound on |value| over the frame
    double grad;         // |(b, g)|: pixels^-1 scale of the strip
};

Affine make_affine(d3 N, double r_float_path, const FrameGeom& g)
{
    Affine A;
    A.a = ddot(N, g.F);
    A.b = g.sp_u * ddot(N, g.right);
    A.g = g.sp_v * ddot(N, g.up);
    A.absmax = std::fabs(A.a) + std::fabs(A.b) * g.xmax + std::fabs(A.g) * g.ymax;
    // (1) float evaluation fma(b, x, fma(g, y, a)) with coefficients rounded to float
    double e_aff = 3.0 * EPS * A.absmax;
    // (2) the reference's un-normalised direction differs from d1(x, y) by <= delta (absolute),
    //     and the two normalisations add a component-wise relative error <= 8 eps
    double Gabs[3] = { g.delta[0] + 8.0 * EPS * g.G[0], g.delta[1] + 8.0 * EPS * g.G[1], g.delta[2] + 8.0 * EPS * g.G[2] };
    double e_dir = absdot(N, Gabs);
    // (3) rounding inside the reference's float cross/dot sequence, scaled back by 1/kappa <= Kmax
    A.T = e_aff + e_dir + r_float_path * g.Kmax;
    A.grad = std::sqrt(A.b * A.b + A.g * A.g);
    return A;
}

// margin of the strip test |X| <= |D| + m
double strip_margin(const Affine& X, const Affine& Dn)
{
    double m = X.T + (1.0 + 8.0 * EPS) * Dn.T + 8.0 * EPS * Dn.absmax + 4.0 * EPS * (Dn.absmax + X.absmax);
    return SAFETY * m + TINY;
}

inline bool finite3(f3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
inline double maxabs(f3 a) { return std::max(std::fabs((double)a.x), std::max(std::fabs((double)a.y), std::fabs((double)a.z))); }

} // namespace

void compute_normal(const float* v, float n[3])
{
    f3 p1 = { v[0], v[1], v[2] }, p2 = { v[3], v[4], v[5] }, p3 = { v[6], v[7], v[8] };
    f3 c = cross(sub(p2, p1), sub(p3, p1));
    float len = std::sqrt(c.x * c.x + c.y * c.y + c.z * c.z);      // Vec3::getLength
    n[0] = c.x / len; n[1] = c.y / len; n[2] = c.z / len;          // Vec3::normalise: three divides
}

FrameGeom make_frame_geom(const float origin[3], const float detector[3], const float up[3], const float right[3],
                          float sp_u, float sp_v, uint32_t W, uint32_t H)
{
    FrameGeom g{};
    g.sp_u = sp_u; g.sp_v = sp_v;
    g.xmax = 0.5 * (double)W; g.ymax = 0.5 * (double)H;
    const double U = std::fabs(g.sp_u) * g.xmax, V = std::fabs(g.sp_v) * g.ymax;
    bool ok = std::isfinite(g.sp_u) && std::isfinite(g.sp_v);
    double G2 = 0;
    for (int i = 0; i < 3; ++i) {
        g.origin[i] = origin[i]; g.up[i] = up[i]; g.right[i] = right[i];
        g.F[i] = (double)detector[i] - (double)origin[i];
        ok = ok && std::isfinite(g.origin[i]) && std::isfinite(g.F[i]) && std::isfinite(g.up[i]) && std::isfinite(g.right[i]);
        // ((detector + up*v) + right*u) - origin in float: <= 5 eps (...) to first order; 6 eps with slack
        g.delta[i] = 6.0 * EPS * (std::fabs((double)detector[i]) + std::fabs(g.up[i]) * V + std::fabs(g.right[i]) * U +
                                  std::fabs(g.origin[i])) + 1e-40;
        g.G[i] = std::fabs(g.F[i]) + std::fabs(g.up[i]) * V + std::fabs(g.right[i]) * U + g.delta[i];
        G2 += g.G[i] * g.G[i];
    }
    g.Kmax = std::sqrt(G2) * (1.0 + 16.0 * EPS);
    // sanity preconditions: bounded magnitudes, detector plane not through the source
    d3 ru = { g.right[0] * g.sp_u, g.right[1] * g.sp_u, g.right[2] * g.sp_u };
    d3 uv = { g.up[0] * g.sp_v, g.up[1] * g.sp_v, g.up[2] * g.sp_v };
    d3 n = dcross(ru, uv);
    double nn = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    double plane_dist = nn > 0 ? std::fabs(ddot(n, g.F)) / nn : 0.0;
    ok = ok && std::isfinite(g.Kmax) && g.Kmax < BIG && plane_dist > 1.0 / 1024.0 && W > 0 && H > 0;
    g.filterable = ok;
    return g;
}

void prepack_mesh(const float* vertices, const float* normals, uint32_t n, const float origin_[3],
                  const FrameGeom& g, uint32_t tile_tris, uint32_t pad_to, PackedMesh& out)
{
    out.n = n;
    out.n_padded = (n + pad_to - 1) / pad_to * pad_to;
    out.exact.resize(n);
    out.filt.resize(out.n_padded);
    out.filt2.resize(out.n_padded);
    out.tile_m.assign((out.n_padded + tile_tris - 1) / tile_tris + 1, 0.0f);
    out.unfilterable = 0;
    const f3 o = { origin_[0], origin_[1], origin_[2] };

    for (uint32_t i = 0; i < n; ++i) {
        const float* v = vertices + 9 * (size_t)i;
        f3 p1 = { v[0], v[1], v[2] }, p2 = { v[3], v[4], v[5] }, p3 = { v[6], v[7], v[8] };
        f3 e1 = sub(p2, p1);                 // Ray.cxx:86
        f3 e2 = sub(p3, p1);                 // Ray.cxx:87
        f3 tv = sub(o, p1);                  // Ray.cxx:102
        f3 qv = cross(tv, e1);               // Ray.cxx:112
        float e2q = dot(e2, qv);             // Ray.cxx:122 (numerator of t)
        float nrm[3];
        if (normals) { nrm[0] = normals[3 * (size_t)i]; nrm[1] = normals[3 * (size_t)i + 1]; nrm[2] = normals[3 * (size_t)i + 2]; }
        else compute_normal(v, nrm);
        ExactTri& E = out.exact[i];
        E.e2x = e2.x; E.e2y = e2.y; E.e2z = e2.z; E.e1x = e1.x;
        E.e1y = e1.y; E.e1z = e1.z; E.tx = tv.x; E.ty = tv.y;
        E.tz = tv.z; E.qx = qv.x; E.qy = qv.y; E.qz = qv.z;
        E.e2q = e2q; E.nx = nrm[0]; E.ny = nrm[1]; E.nz = nrm[2];

        FilterTri& f = out.filt[i];
        FilterTri2& f2 = out.filt2[i];
        bool ok = g.filterable && finite3(e1) && finite3(e2) && finite3(tv) && finite3(qv) &&
                  maxabs(e1) < BIG && maxabs(e2) < BIG && maxabs(tv) < BIG && maxabs(qv) < BIG * BIG;
        if (ok) {
            d3 E1 = D(e1), E2 = D(e2), TV = D(tv), QV = D(qv);
            // s = tvec . (d x e2) = d . (e2 x tvec);  det = e1 . (d x e2) = d . (e2 x e1)
            d3 Ns = dcross(E2, TV), Nd = dcross(E2, E1);
            auto ab = [](d3 a) { return d3{ std::fabs(a.x), std::fabs(a.y), std::fabs(a.z) }; };
            d3 a1 = ab(E1), a2 = ab(E2), at = ab(TV), aq = ab(QV);
            double A = at.x * (a2.y + a2.z) + at.y * (a2.x + a2.z) + at.z * (a2.x + a2.y);
            double B = a1.x * (a2.y + a2.z) + a1.y * (a2.x + a2.z) + a1.z * (a2.x + a2.y);
            double Cq = aq.x + aq.y + aq.z;
            double rS = 6.0 * EPS * A + TINY, rD = 6.0 * EPS * B + TINY, rV = 4.0 * EPS * Cq + TINY;
            Affine S = make_affine(Ns, rS, g);
            Affine Dn = make_affine(Nd, rD, g);
            Affine Vn = make_affine(QV, rV, g);
            d3 Nw = { Nd.x - Ns.x - QV.x, Nd.y - Ns.y - QV.y, Nd.z - Ns.z - QV.z };
            Affine Wn = make_affine(Nw, rS + rD + rV, g);
            // the w strip compares against |D|(1 + 8 eps): already inside strip_margin's 8 eps term
            Affine* st[3] = { &S, &Vn, &Wn };
            // narrowest strip first (largest gradient => fewest surviving pixels)
            std::sort(st, st + 3, [](const Affine* x, const Affine* y) { return x->grad > y->grad; });
            double m0 = strip_margin(*st[0], Dn), m1 = strip_margin(*st[1], Dn), m2 = strip_margin(*st[2], Dn);
            double lim = 1e30;
            ok = std::isfinite(m0) && std::isfinite(m1) && std::isfinite(m2) && m0 < lim && m1 < lim && m2 < lim &&
                 S.absmax < lim && Dn.absmax < lim && Vn.absmax < lim && Wn.absmax < lim;
            if (ok) {
                f.bx = (float)st[0]->b; f.gx = (float)st[0]->g; f.ax = (float)st[0]->a;
                f.bd = (float)Dn.b; f.gd = (float)Dn.g; f.ad = (float)Dn.a;
                f.m = round_up(m0); f.pad = 0.0f;
                f2.bx2 = (float)st[1]->b; f2.gx2 = (float)st[1]->g; f2.ax2 = (float)st[1]->a; f2.m2 = round_up(m1);
                f2.bx3 = (float)st[2]->b; f2.gx3 = (float)st[2]->g; f2.ax3 = (float)st[2]->a; f2.m3 = round_up(m2);
            }
        }
        if (!ok) {
            // always survives: X == D == 0 -> |D| - |X| = 0 >= -m
            f = FilterTri{ 0, 0, 0, 0, 0, 0, 0, 0 };
            f2 = FilterTri2{ 0, 0, 0, 0, 0, 0, 0, 0 };
            ++out.unfilterable;
        }
        float& tm = out.tile_m[i / tile_tris];
        tm = std::max(tm, f.m);
    }
    // padding entries never survive: X = 1e30, D = 0 -> |D| - |X| = -1e30 < -m for every tile margin (< 1e30)
    for (uint32_t i = n; i < out.n_padded; ++i) {
        out.filt[i] = FilterTri{ 0, 0, 3.0e38f, 0, 0, 0, 0, 0 };
        out.filt2[i] = FilterTri2{ 0, 0, 3.0e38f, 0, 0, 0, 3.0e38f, 0 };
    }
}

} // namespace srt
