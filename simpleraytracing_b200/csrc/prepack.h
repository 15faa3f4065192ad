// prepack.h -- host-side re-pack of a TriangleMesh (diagnostic path) and the per-camera frame geometry.
//
// The reference stores Triangle as 7 x Vec3 AoS (include/Triangle.h:125-133) and recomputes edge
// vectors, tvec and qvec inside every Ray::intersect call (src/Ray.cxx:86-122).  All primary rays
// share one origin (src/main-pthreads-lbuffer.cxx:762), so these are per-triangle constants.  The
// render path computes them on the GPU (prepack kernel in kernels.cu); the same arithmetic
// (prepack_math.h) is available on the host for srt_debug_prepack and the CPU filter audit.
#pragma once
#include <cstdint>
#include <vector>

#include "prepack_math.h"

namespace srt {

// box_w x box_h: the box of pixels one thread of the filtered kernel tests per triangle (FilterTri::R is packed for it)
FrameGeom make_frame_geom(const float origin[3], const float detector[3], const float up[3], const float right[3],
                          float sp_u, float sp_v, uint32_t W, uint32_t H, uint32_t box_w, uint32_t box_h);

struct PackedMesh {
    std::vector<ExactTri>   exact;    // n
    std::vector<FilterTri>  filt;     // n padded to a multiple of `pad_to` with never-surviving entries
    std::vector<FilterTri2> filt2;    // n padded likewise
    uint32_t n = 0, n_padded = 0;
    uint32_t unfilterable = 0;        // triangles that always take the exact path
};

// vertices: 9 floats per triangle; normals: 3 per triangle or nullptr.
void prepack_mesh(const float* vertices, const float* normals, uint32_t n, const float origin[3],
                  const FrameGeom& g, uint32_t pad_to, PackedMesh& out);

} // namespace srt
