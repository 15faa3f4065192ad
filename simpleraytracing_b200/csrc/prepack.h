// prepack.h -- host-side re-pack of a TriangleMesh for the sm_100a kernels.
//
// The reference stores Triangle as 7 x Vec3 AoS (include/Triangle.h:125-133) and
// recomputes edge vectors, tvec and qvec inside every Ray::intersect call
// (src/Ray.cxx:86-122).  All primary rays share one origin
// (src/main-pthreads-lbuffer.cxx:762), so these are per-triangle constants; they
// are computed here ONCE with the reference's exact float operation order (this
// translation unit is compiled with -ffp-contract=off) and streamed to the GPU
// as float4 tiles.
#pragma once
#include <cstdint>
#include <vector>

namespace srt {

// 64 B: everything the exact Moller-Trumbore sequence needs (4 x float4).
struct ExactTri {
    float e2x, e2y, e2z, e1x;
    float e1y, e1z, tx, ty;
    float tz, qx, qy, qz;
    float e2q, nx, ny, nz;       // e2q = edge2 . qvec (Ray.cxx:122 numerator), n = Triangle::getNormal()
};

// 32 B: level-0 conservative reject.  X and D are affine in the centred pixel
// coordinates (x, y): X = bx*x + gx*y + ax, D = bd*x + gd*y + ad.  A pair can
// only be a hit when |X| <= |D| + m (see DESIGN.md "filter proof").
struct FilterTri {
    float bx, gx, ax, bd;
    float gd, ad, m, pad;
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
    bool   filterable;           // camera passes the sanity preconditions of the proof
};

FrameGeom make_frame_geom(const float origin[3], const float detector[3], const float up[3], const float right[3],
                          float sp_u, float sp_v, uint32_t W, uint32_t H);

// Triangle::computeNormal (include/Triangle.inl:170-178), exact op order.
void compute_normal(const float* v9, float n[3]);

struct PackedMesh {
    std::vector<ExactTri>   exact;    // n
    std::vector<FilterTri>  filt;     // n padded to a multiple of `pad_to` with never-surviving entries
    std::vector<FilterTri2> filt2;    // n padded likewise
    std::vector<float>      tile_m;   // max level-0 margin per tile of `tile_tris` triangles
    uint32_t n = 0, n_padded = 0;
    uint32_t unfilterable = 0;        // triangles that always take the exact path
};

// vertices: 9 floats per triangle; normals: 3 per triangle or nullptr.
void prepack_mesh(const float* vertices, const float* normals, uint32_t n, const float origin[3],
                  const FrameGeom& g, uint32_t tile_tris, uint32_t pad_to, PackedMesh& out);

} // namespace srt
