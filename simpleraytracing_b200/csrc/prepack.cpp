// prepack.cpp -- see prepack.h.  Compile with -ffp-contract=off (no FMA contraction): the exact float
// results must equal the reference's (src/Ray.cxx:86-122, include/Triangle.inl:170-178) bit for bit.
#include "prepack.h"

#include <algorithm>
#include <cmath>

#if defined(__FMA__) && !defined(SRT_ALLOW_FMA_HOST)
#error "build prepack.cpp without -mfma/-march=native (contraction would change the bits)"
#endif

namespace srt {

FrameGeom make_frame_geom(const float origin[3], const float detector[3], const float up[3], const float right[3],
                          float sp_u, float sp_v, uint32_t W, uint32_t H, uint32_t box_w, uint32_t box_h)
{
    using namespace pm;
    FrameGeom g{};
    g.sp_u = sp_u; g.sp_v = sp_v;
    g.box_hx = 0.5 * (double)(box_w ? box_w - 1 : 0); g.box_hy = 0.5 * (double)(box_h ? box_h - 1 : 0);
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
    const d3 ru = { g.right[0] * g.sp_u, g.right[1] * g.sp_u, g.right[2] * g.sp_u };
    const d3 uv = { g.up[0] * g.sp_v, g.up[1] * g.sp_v, g.up[2] * g.sp_v };
    const d3 n = dcross(ru, uv);
    const double nn = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    const double plane_dist = nn > 0 ? std::fabs(ddot(n, g.F)) / nn : 0.0;
    ok = ok && std::isfinite(g.Kmax) && g.Kmax < BIG && plane_dist > 1.0 / 1024.0 && W > 0 && H > 0;
    g.filterable = ok ? 1 : 0;
    return g;
}

void prepack_mesh(const float* vertices, const float* normals, uint32_t n, const float origin[3],
                  const FrameGeom& g, uint32_t pad_to, PackedMesh& out)
{
    out.n = n;
    out.n_padded = (n + pad_to - 1) / pad_to * pad_to;
    out.exact.resize(n);
    out.filt.resize(out.n_padded);
    out.filt2.resize(out.n_padded);
    out.unfilterable = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const bool ok = pack_triangle(vertices + 9 * (size_t)i, normals ? normals + 3 * (size_t)i : nullptr, origin, g,
                                      out.exact[i], out.filt[i], out.filt2[i]);
        if (!ok) ++out.unfilterable;
    }
    for (uint32_t i = n; i < out.n_padded; ++i) pack_padding(out.filt[i], out.filt2[i]);
}

} // namespace srt
