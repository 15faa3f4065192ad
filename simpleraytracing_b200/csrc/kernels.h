// kernels.h -- launch interface between the C-ABI layer (srt_api.cpp) and the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace srt {

constexpr uint32_t kPadTris = 4;       // meshes are padded to a multiple (never-surviving records)

// One mesh as the kernels see it (device pointers).
struct DevMesh {
    const float4* filt0;     // 3 float4 per (padded) triangle: FilterTri  (level 0), natural triangle order
    const float4* scan;      // the same records, chunk-major SoA for the filtered kernel: [n_chunks][scan_k][3][128]
    uint32_t n_chunks;       // chunks of 128 * scan_k triangles (a strided sample of the mesh each)
    const float4* filt1;     // 2 float4 per (padded) triangle: FilterTri2 (level 1: the two other strips)
    const float4* exact;     // 4 float4 per triangle            -- ExactTri
    uint32_t n, n_padded;
    float    mu;             // attenuation coefficient
    int      contributes;    // 0: mesh adds no path length (reference semantics for meshes > 0)
    uint32_t first;          // scene-wide index of the mesh's first triangle (closest-hit mode)
};

struct DevCamera {
    float ox, oy, oz;        // origin
    float dx, dy, dz;        // detector
    float ux, uy, uz;        // up
    float rx, ry, rz;        // right
    const float* u_off;      // W entries: float(double(spacing)*(0.5 + col - W/2.0)), lbuffer.cxx:757
    const float* v_off;      // H entries, lbuffer.cxx:756
    float xc, yc;            // centred pixel coordinate = index + xc (0.5 - W/2)
    uint32_t W, H;
};

struct RenderArgs {
    DevCamera cam;
    const DevMesh* meshes;   // device array
    int n_meshes;
    uint32_t row_begin, row_end;
    // Row-block interleaving (multi-GPU load balance): CTA row-block `by` of the grid renders the image rows of
    // block `by * blk_stride + blk_phase` (16 rows each, counted from row_begin).  blk_stride = 1, blk_phase = 0
    // is a contiguous band.  compact != 0: outputs are addressed by the LOCAL row `by * 16 + r` (a dense slot
    // buffer for the all_gather) instead of the image row.
    uint32_t blk_stride, blk_phase, n_blocks;   // n_blocks = grid.y
    int compact;
    float incident;          // 80.0f
    double length_to_cm;     // 0.1
    float* L;                // full-frame outputs (row*W + col); band rows written
    // Fused assembly (multi-GPU): with L_image != 0 the L value of a compact (block-cyclic) launch is written at its
    // IMAGE position instead of the slot position, and additionally into n_L_extra more full-frame buffers -- the
    // peers' frames mapped over NVLink, or one multicast mapping that the switch replicates to every GPU.
    int L_image;
    int L_vec4;              // every L frame is 16-byte aligned and W % 4 == 0: 128-bit stores allowed
    int L_multicast;         // L is an NVSwitch multicast mapping: stored with multimem.st (plain st to such an address is undefined)
    uint32_t n_L_extra;
    float* L_extra[7];
    int L_halo;              // > 0: L_extra[0] / L_extra[1] are the frames of the ranks owning the blocks above / below this rank's;
                             // only the first / last L_halo rows of each 16-row block are stored there (the repair stencil's reach)
    float* dist;             // optional
    int32_t* ssum;           // optional
    uint32_t* hits;          // optional
    unsigned long long* counters;   // optional: [0] exact tests run, [1] filter violations (audit), [2] two floats: the smallest
                                    // head-room of the strip bounds and of the level-0 box bounds over all hits (audit)
    // mode 0: X-ray L-buffer (above).  mode 1: closest hit over all meshes (shaded variant, main-omp.cxx:632-666):
    // `dist` receives the z-buffer value t (+inf = no hit), `ssum` the scene-wide triangle index (-1 = none);
    // strict `z > t` in ascending triangle order => the first of equal t wins, negative t is accepted.
    // mode 2 (filtered path only): sort-based X-ray of src/main.cxx:626-743: the ray's t > 1e-7 of the contributing meshes
    // are sorted and exit - entry pairs summed; `L` receives 80 exp(-(mu0 d / 10)), `dist` the path length d, `hits` the count.
    int mode;
    // Filtered kernel.  A thread holds scan_k triangles and tests them against boxes of box_px x box_py pixels
    // (FilterTri::R / R2 are packed for that box); the unit of work is one box row of a 16-row block x scan_cols box
    // columns x one chunk of 128 * scan_k triangles.
    // Accepted intersections are not accumulated in place (the reference's float sum is order dependent and the exact
    // tests of a warp run wherever a lane is free) but appended to the pixel's hit list {scene-wide triangle index,
    // value} -- value = sign * t (mode 0) or t (modes 1, 2) -- a linked list with an atomic head; resolve_hits_kernel
    // sorts each list by triangle index and replays lbuffer.cxx:788-809 (or the z-buffer update, main-omp.cxx:656-662) in the
    // reference's order.
    uint32_t box_px, box_py, scan_k, scan_cols, scan_chunks;
    // the scan runs scan_ctas persistent CTAs over the units (chunk, block, box row, range of scan_cols box columns):
    // scan_chunks = chunks of all contributing meshes together, scan_parts_x = column ranges per box row
    uint32_t scan_parts_x, scan_ctas;
    uint32_t scan_blk_stride;    // the t-th block a CTA visits within a chunk is block (t * scan_blk_stride) % n_blocks (coprime stride)
    uint32_t sm_count;
    // survivors of the scan {box, triangle} (cand) and of the per-pixel strip tests {pixel, triangle} (pix).
    // cand_count[0] / [1] count them: zero when the scan starts; they may exceed the capacities (entries beyond are
    // dropped and resolve_hits_kernel recomputes every pixel with the exact sequence).
    uint2* cand; uint2* pix; uint32_t* cand_count; uint32_t cand_cap, pix_cap;
    // Big-box shapes (box_px * box_py > 64 pixels): the scan's survivors {big box, triangle} go to `big` (counted in
    // *big_count); refine_kernel tests the 16 x 4 sub-boxes of every entry with the three-strip box test and appends the
    // surviving {sub-box, triangle} to `cand`, which candidates_kernel<16, 4> consumes as for the small shapes.
    uint2* big; uint32_t* big_count; uint32_t big_cap, big_hint;
    uint32_t* cand_count_next;   // the other set of the double-buffered counters (4 words: cand, pix, big, -): the scan leaves it at zero for the next launch (may be null)
    uint32_t cand_hint, pix_hint;   // list sizes of this launch geometry when the host has verified them (0 = unknown): grid sizing only
    uint32_t scan_max_ctas;      // cap on the scan's resident CTAs per SM (0 = none), see scan_smem_bytes
    uint32_t* hit_head;      // [pixel]: the pixel's hit list (index into pix / hit_next; 0xffffffff = empty, also between renders)
    uint32_t* hit_next;      // [pix_cap]: next node of the list; node e = candidate e of the pix list, which exact_tests_kernel
                             // overwrites in place with {scene-wide triangle index, value}
    // Culled kernel (SRT_KERNEL_CULLED): per-bin triangle lists.  A bin is kBinCols columns x bin_blocks 16-row blocks
    // of this launch's grid; bin_entries[bin * bin_cap + off_m + i], i < bin_counts[bin * n_meshes + m] (off_m: the mesh's
    // proportional share of the bin_cap entries of a bin; a count above the share = "walk the whole mesh"), are the
    // indices (within mesh m, ascending) of the triangles whose conservative footprint may touch the bin.
    uint32_t* bin_entries;
    uint32_t* bin_counts;
    uint32_t bin_cap, bin_blocks, bins_x, bins_y, total_tris;
};
constexpr uint32_t kBinCols = 256;     // bin width in pixels (4 CTA tiles of the culled kernel)
cudaError_t launch_resolve_hits(const RenderArgs& a, cudaStream_t s);

// ---- shaded variant: Phong + shadow ray + texture for the pixels of a closest-hit pass ------------------
struct ShadeMesh {
    const float*   vertices;   // 9 floats per triangle
    const float4*  exact;      // ExactTri (for the cached normals)
    const float*   texcoords;  // 9 floats per triangle or nullptr
    const uint8_t* tex;        // RGB texture or nullptr
    uint32_t tex_w, tex_h;
    uint32_t n, first;
    float ka[3], kd[3], ks[3], shininess;
};

// Light-space grid for the shadow rays of the shaded variant.  Every shadow ray lies on a line through the light, so a
// triangle can only block the ray of hit point P if its central projection from the light onto a plane (basis u, v,
// axis w) covers P's projection.  Triangles are binned by the bounding box of their projected vertices, widened by one
// cell on every side (float rounding of the reference's test moves a ray by ~1e-7 of its length; a cell is ~2e-3 of the
// scene's extent); triangles that cross the plane through the light (no finite projection) or cover more than
// kShadowMaxCells cells go to a short list every ray tests.  Each ray then runs the reference's exact test only on the
// triangles of its cell: the boolean "any blocker" does not depend on the order.
struct ShadowGrid {
    int enabled;               // 0: brute force over all triangles (light inside the scene's bounding box, ...)
    float u[3], v[3], w[3];    // orthonormal light-space basis
    float a0, b0, inv_da, inv_db;   // grid origin and 1 / cell size in projected coordinates
    uint32_t G;                // cells per side
    uint32_t* cell_off;        // [G*G + 1] exclusive prefix of the cell counts (count pass: the counts)
    uint32_t* cell_cursor;     // [G*G] fill pass: next free slot of each cell
    uint32_t* cell_tris;       // scene-wide triangle indices, cell by cell
    uint32_t* global_tris;     // triangles every ray tests
    uint32_t* n_global;
    uint32_t total_tris;
};
constexpr uint32_t kShadowMaxCells = 1024;

struct ShadeArgs {
    DevCamera cam;
    const ShadeMesh* meshes;   // device array
    int n_meshes;
    float light_colour[3], light_position[3];
    const float* z;            // closest-hit pass outputs (full frame)
    const int32_t* tri_id;
    uint8_t* rgb;              // full frame, 3 bytes per pixel
    uint8_t* shadow;           // optional: 1 where the shadow ray is blocked
    uint32_t row_begin, row_end;
    uint8_t background[3];
    ShadowGrid grid;
};
cudaError_t launch_shade(const ShadeArgs& a, cudaStream_t s);
// bounding box of all vertices of the scene: out6 = {min x, y, z, max x, y, z} as order-preserving unsigned keys
// (decode with ordered_key_to_float); out6 must be preset to {0xffffffff x 3, 0 x 3}
cudaError_t launch_scene_bbox(const ShadeMesh* meshes, int n_meshes, uint32_t total_tris, uint32_t* out6, cudaStream_t s);
float ordered_key_to_float(uint32_t k);
// shadow grid: count pass (fill = 0: cell_off receives the counts), in-place exclusive scan of cell_off[0 .. G*G],
// fill pass (fill = 1: cell_cursor preset to cell_off)
cudaError_t launch_shadow_bin(const ShadeArgs& a, int fill, cudaStream_t s);
cudaError_t launch_exclusive_scan(uint32_t* data, uint32_t n, cudaStream_t s);

// Exact kernel: every pair runs the reference op sequence. audit != 0 additionally evaluates the
// filters for every pair and counts (filter says reject) && (exact says hit) into counters[1].
cudaError_t launch_lbuffer_exact(const RenderArgs& a, int audit, cudaStream_t s);
// Filtered path for the shape (a.box_px, a.box_py, a.scan_k): the scan (a.cand_count[0..1] must be zero), then the
// candidate stages (per-pixel strip tests, exact tests), then launch_resolve_hits.
cudaError_t launch_lbuffer_filtered(const RenderArgs& a, cudaStream_t s);
cudaError_t launch_filtered_candidates(const RenderArgs& a, cudaStream_t s);
bool filtered_shape_supported(uint32_t box_px, uint32_t box_py, uint32_t scan_k);
int filtered_resident_ctas(uint32_t box_px, uint32_t box_py, uint32_t scan_k, uint32_t max_ctas = 0);     // CTAs resident per SM on the current device
// Culled kernel: bin pass (every triangle against every bin of a.bins_x x a.bins_y), then the render pass.  Same results
// as the other kernels: the bins and the per-CTA cull only ever drop pairs the conservative filter proves to miss.
cudaError_t launch_lbuffer_bins(const RenderArgs& a, cudaStream_t s);
cudaError_t launch_lbuffer_culled(const RenderArgs& a, cudaStream_t s);
// Scatter gathered slots [n_parts][slot_rows][W] (block-cyclic 16-row blocks) back to image order [H][W].
cudaError_t launch_deinterleave(const float* slots, float* frame, uint32_t W, uint32_t H, uint32_t n_parts,
                                uint32_t slot_rows, cudaStream_t s);
// Device-side barrier over the n ranks of a node: p[j] = rank j's flag array (n words, zero before the first epoch) as
// mapped in this process; epochs start at 1 and grow by one per barrier.
struct PeerFlags { uint32_t* p[8]; };
// phases: 1 = signal only, 2 = wait only, 3 = both (a caller that drives several ranks from one thread enqueues all the
// signals before the first wait)
// epoch_ctr != nullptr: the barrier's epoch is ++(*epoch_ctr) (device memory) instead of `epoch`
cudaError_t launch_peer_barrier(const PeerFlags& f, uint32_t n, uint32_t rank, uint32_t epoch, uint32_t phases, uint32_t* epoch_ctr,
                                cudaStream_t s);
// Artefact repair + image fill (lbuffer.cxx:329-404).
cudaError_t launch_repair(const float* L, float* image, uint32_t W, uint32_t H, uint32_t depth, cudaStream_t s);
// Sharded repair: the pixels of the 16-row blocks part, part + n_parts, ... of the frame, stored into `image` (a multicast
// mapping when multicast != 0) and the n_extra extra images; vec4: every image is 16-byte aligned and W % 4 == 0.
struct RepairOut { float* image; int multicast, vec4; uint32_t n_extra; float* extra[7]; };
cudaError_t launch_repair_blocks(const float* L, const RepairOut& out, uint32_t W, uint32_t H, uint32_t depth, uint32_t n_parts,
                                 uint32_t part, cudaStream_t s);

// Scene pre-pack on the GPU (prepack_math.h): raw vertices -> ExactTri / FilterTri / FilterTri2.
struct FrameGeom;
cudaError_t launch_prepack(const float* vertices, const float* normals, uint32_t n, uint32_t n_padded, uint32_t n_chunks,
                           uint32_t scan_k, const float origin[3], const FrameGeom& g, void* exact, void* filt0, void* filt1,
                           void* scan, cudaStream_t s);
// Detector offset tables u_off[W], v_off[H] (lbuffer.cxx:756-757).
cudaError_t launch_offsets(float* u_off, float* v_off, uint32_t W, uint32_t H, float sp_u, float sp_v, cudaStream_t s);

// FP32 FMA throughput probe: blocks x 256 threads x iters x 16 FFMA.
cudaError_t launch_ffma_peak(float* scratch, int blocks, int iters, cudaStream_t s);

} // namespace srt
