/*
 * srt.h -- C ABI of the B200-native SimpleRayTracing hot path.
 *
 * The reference (Brandagot/SimpleRayTracing) has no plugin/FFI layer: every
 * driver inlines its own per-pixel loop.  The replaceable unit is the
 * driver-level render function, and this header is the boundary a maintainer
 * binds instead of it (see INTEGRATION.md for the C++ shim):
 *
 *   reference interface                                     replaced by
 *   ------------------------------------------------------  ---------------------------
 *   renderLoop(Image&, const vector<TriangleMesh>&,         srt_set_meshes + srt_set_camera
 *              RayTracerInfo&)            src/main.cxx:159-161,626      + srt_render_xray_sorted (the driver's own
 *                                                                       sort-based image) or srt_render_xray (the
 *                                                                       L-buffer image of main-pthreads-lbuffer)
 *   void* renderLoopCallBack(void* PThreadData*)            srt_render_lbuffer (row band)
 *              src/main-pthreads-lbuffer.cxx:186,733-813
 *   global std::vector<float> L_buffer    ...lbuffer.cxx:254,314        L_out of srt_render_lbuffer
 *   repair + Image::setPixel loop         ...lbuffer.cxx:329-404        srt_repair
 *   initialiseRayTracing / getBBox        ...lbuffer.cxx:585-667        srt_scene_bbox + srt_camera_from_bbox
 *   pthreadWorkLoadAllocation             ...lbuffer.cxx:694-725        srt_partition_rows
 *   MPI gather of pixel ranges            src/main-mpi.cxx:854-881      srt_multi_* (one process, N GPUs) or, one
 *                                                                       process per GPU, srt_render_xray_blocks_device
 *                                                                       on peer-mapped buffers: the kernels store the
 *                                                                       rows other GPUs need over NVLink, device-side
 *                                                                       barriers in between
 *   renderLoop(Image&, meshes, detector, origin, up, right, Light)      srt_render_shaded
 *              src/main-omp.cxx:136-142,588-797
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success
 * and a negative srt_status on failure (message: srt_last_error); nothing
 * throws across the boundary; the caller owns every host pointer, the library
 * owns all device memory except the buffers handed to the *_device entry
 * points.  One context drives one GPU (one process per GPU); calls on a
 * context must not be made concurrently.  There is no CPU fallback: without a
 * usable sm_100 device srt_create fails.
 */
#ifndef SRT_H
#define SRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRT_VERSION 100

typedef enum {
    SRT_OK = 0,
    SRT_ERR_INVALID = -1,      /* bad argument (NULL, sizes, order of calls)  ~ std::length_error / out_of_range */
    SRT_ERR_CUDA = -2,         /* CUDA runtime failure                                                */
    SRT_ERR_NO_DEVICE = -3,    /* no CUDA device / not an sm_100 part                                 */
    SRT_ERR_NOMEM = -4
} srt_status;

typedef struct srt_ctx srt_ctx;

/* Material (include/Material.h:86-124): plain data, shaded variant only. */
typedef struct {
    float ambient[3], diffuse[3], specular[3];
    float shininess;
} srt_material;

/* One TriangleMesh (include/TriangleMesh.h:106-169) flattened. */
typedef struct {
    const float*   vertices;    /* 9 floats per triangle: P1 P2 P3 (Triangle::getP1..P3)             */
    const float*   normals;     /* 3 per triangle == Triangle::getNormal(); NULL -> computed exactly
                                   as Triangle::computeNormal (include/Triangle.inl:170-178)          */
    const float*   texcoords;   /* 9 per triangle (3 x Vec3) or NULL                                  */
    uint32_t       n_triangles;
    float          mu;          /* linear attenuation coefficient of this mesh's material (1/cm);
                                   the reference hard-codes 0.1037f (mesh 0) / 0.3971f (others),
                                   src/main-pthreads-lbuffer.cxx:798-803                               */
    srt_material   material;
    const uint8_t* texture_rgb; /* tex_w*tex_h*3, row-major, or NULL                                  */
    uint32_t       tex_w, tex_h;
} srt_mesh_desc;

/* RayTracerInfo (src/main-pthreads-lbuffer.cxx:125-135) + image geometry. */
typedef struct {
    float    origin[3];         /* X-ray source == ray origin of every primary ray                    */
    float    detector[3];       /* detector centre                                                    */
    float    up[3], right[3];
    float    pixel_spacing[2];  /* [0]: columns (u), [1]: rows (v); ...lbuffer.cxx:739-741            */
    uint32_t width, height;
} srt_camera;

/* Light (include/Light.h:80-111); the shaded drivers use a white light at origin + up * 100.0
 * (src/main-omp.cxx:235-238).  Only colour and position enter the shading (applyShading, :407-434). */
typedef struct {
    float colour[3];
    float direction[3];
    float position[3];
} srt_light;

typedef enum {
    SRT_KERNEL_AUTO = 0,        /* fastest for the scene: exact for a handful of triangles, filtered (brute force)
                                   up to ~2e5 triangles, culled beyond (same results whichever runs)      */
    SRT_KERNEL_EXACT = 1,       /* every (ray, triangle) pair runs the reference op sequence          */
    SRT_KERNEL_FILTERED = 2,    /* conservative affine reject per pair + exact test for survivors     */
    SRT_KERNEL_AUDIT = 3,       /* diagnostic: exact kernel that also evaluates the filter for every
                                   pair and counts (filter rejects && exact hits) -- must stay 0      */
    SRT_KERNEL_CULLED = 4       /* the same conservative filter applied hierarchically (image bins, CTA
                                   tiles, pixel runs) before the per-pair tests: identical hit set and
                                   summation order, the accelerated path the LBVH option stands for   */
} srt_kernel;

#define SRT_FLAG_ALL_MESHES_CONTRIBUTE 1u  /* default off: only mesh 0 adds path length (lbuffer.cxx:788) */
#define SRT_FLAG_FRAME0_MULTICAST 2u       /* sharded entry points: a frame / image array of ONE pointer is an NVSwitch
                                              multicast mapping of every GPU's buffer (written with multimem.st)   */
#define SRT_FLAG_L_HALO_ONLY 4u            /* srt_render_lbuffer_blocks_device with n_frames == n_parts (own frame, then
                                              every peer's in rank order): store each L row into the own frame and
                                              only the first / last repair_depth rows of a 16-row block into the frame
                                              of the rank that owns the block above / below -- all the sharded repair
                                              pass (srt_repair_blocks_device) reads; 1/16 of the NVLink traffic of
                                              replicating every row to every GPU                                   */

typedef struct {
    float    incident;          /* initial L-buffer value, 80.0f (lbuffer.cxx:314)                    */
    double   length_to_cm;      /* 0.1: mesh units (mm) -> cm (lbuffer.cxx:808)                       */
    uint32_t repair_depth;      /* 4 (lbuffer.cxx:339)                                                */
    uint32_t kernel;            /* srt_kernel                                                         */
    uint32_t flags;
    uint32_t pixels_per_thread; /* 0 = auto; tuning knob of the filtered kernel: the box of pixels tested per triangle
                                   and the triangles a thread holds, box width | box height << 8 | triangles << 16
                                   (e.g. 0x040210 = 16 x 2 pixels, 4 triangles)                                  */
} srt_params;

typedef struct {
    float    render_ms;         /* device time of the last render (events on the context's stream)   */
    float    repair_ms;
    float    h2d_ms, d2h_ms;
    uint64_t pair_tests;        /* (ray, triangle) pairs evaluated by the last render                 */
    uint64_t exact_tests;       /* pairs that ran the exact reference sequence (srt_debug_count_exact) */
    uint64_t filter_violations; /* SRT_KERNEL_AUDIT only: pairs the filter would have wrongly rejected */
    uint32_t kernel_used;       /* srt_kernel actually launched                                       */
    uint32_t launches;          /* kernels launched by the last call                                  */
    float    main_kernel_ms;    /* the dominant kernel of the last render alone: lbuffer_filtered_kernel /
                                   lbuffer_culled_kernel / lbuffer_exact_kernel (render_ms also covers the
                                   second kernel of the pass: hit-list resolve, resp. the bin pass)   */
    uint32_t scan_cols;         /* filtered kernel: box columns per CTA of the last render            */
    float    min_strip_slack;   /* SRT_KERNEL_AUDIT: smallest head-room of a strip bound on a HIT, in units of the
                                   bound's margin: (|D| + m - |Y|) / m (3.4e38 if there was no hit)        */
    float    min_box_slack;     /* the same for the level-0 box bounds: (R - (|Y(c)| - |D(c)|)) / 2m; the box radii
                                   in R are exact geometry, so 0.5 is the geometric worst case (a hit on the
                                   strip's edge at a box corner)                                       */
    float    shadow_grid_ms;    /* srt_render_shaded: device + host time spent (re)building the shadow rays' grid
                                   (0 when the grid of the previous call was reused); repair_ms = the shade pass */
    uint32_t shadow_grid_cells; /* cells per side of that grid; 0 = brute force (light inside the scene's box)  */
    uint32_t shadow_grid_entries;
    uint32_t shadow_grid_rebuilt;
} srt_timing;

/* ---- context ------------------------------------------------------------ */
int  srt_create(srt_ctx** out, int device_id);
void srt_destroy(srt_ctx* ctx);
const char* srt_last_error(const srt_ctx* ctx);          /* ctx may be NULL: last create error      */
int  srt_version(void);
/* Launch on an existing CUDA stream (cudaStream_t as void*); NULL = the context's own stream. */
int  srt_set_stream(srt_ctx* ctx, void* cuda_stream);
int  srt_synchronize(srt_ctx* ctx);
void srt_default_params(srt_params* p);
/* Per-render CUDA timing events (srt_last_timing's render_ms / main_kernel_ms / repair_ms): on by default; a caller
 * that pipelines frames on several contexts / streams switches them off (six event records per frame are a
 * measurable share of a 0.1 ms frame).  With profiling off the timing fields read 0. */
int  srt_set_profiling(srt_ctx* ctx, int enable);
/* Cap the resident CTAs per SM of the brute-force scan kernel (0 = no cap, else >= 4): a scan that leaves part of the
 * register file free lets the short, latency-bound kernels of another frame (another context on another stream) run
 * beside it.  Tuning knob for frame pipelining; results never change. */
int  srt_set_scan_residency(srt_ctx* ctx, uint32_t max_ctas_per_sm);
/* Host threads (besides the caller's) that move a frame into a PAGEABLE buffer the caller passed to a host-pointer entry
 * point (srt_render_xray, srt_render_xray_sorted, srt_render_lbuffer, srt_repair, srt_render_shaded,
 * srt_multi_render_xray): frames of 2 MiB or more land chunk by chunk in a pinned staging area of the context and these
 * threads copy every chunk that has arrived while the next ones are in flight (the driver's own pageable copy runs on
 * one thread: 1.0 ms for a 2048^2 frame).  -1 = default (a quarter of the host's cores, at most 4; environment variable
 * SRT_COPY_THREADS overrides), 0 = off (plain cudaMemcpyAsync), at most 32.  Page-locked buffers (srt_host_alloc) are
 * always written by the DMA engine directly.  The bytes returned never depend on this setting. */
int  srt_set_host_copy_threads(srt_ctx* ctx, int threads);

/* ---- host-side helpers (pure CPU arithmetic, exact reference op order) --- */
/* TriangleMesh::computeBoundingBox + getBBox (src/TriangleMesh.cxx:192-228, lbuffer.cxx:585-608). */
int  srt_scene_bbox(const srt_mesh_desc* meshes, int n_meshes, float lower[3], float upper[3]);
/* initialiseRayTracing (lbuffer.cxx:611-667) + pixel spacing (:739-741). */
int  srt_camera_from_bbox(const float lower[3], const float upper[3], uint32_t width, uint32_t height,
                          srt_camera* out);
/* Row-band partition with the reference's remainder rule (main-pthreads-rows.cxx:312-334 /
 * lbuffer.cxx:694-725 applied to rows): part i of n gets rows [begin, end). */
int  srt_partition_rows(uint32_t height, uint32_t n_parts, uint32_t part, uint32_t* row_begin, uint32_t* row_end);
/* Triangle::computeNormal for n triangles (9 floats in, 3 out each). */
int  srt_triangle_normals(const float* vertices, uint32_t n_triangles, float* normals_out);

/* Page-locked host memory for images / vertex arrays the caller transfers every frame (cudaHostAlloc): host-pointer
 * entry points accept any host pointer, pinned ones skip the staging copy of the driver. */
int  srt_host_alloc(void** out, size_t bytes);
void srt_host_free(void* p);

/* JPEG writer (nvJPEG encoder on the context's GPU): 8-bit interleaved RGB, quality 1 ... 100, no chroma subsampling.
 * Replaces the libjpeg writers Image::saveJPEGFile (src/Image.cxx:85-146: quality 100, 3 components) and the old
 * RGB Image's (src/Image.bakk.cxx:170-226). */
int  srt_write_jpeg(srt_ctx* ctx, const char* path, const uint8_t* rgb, uint32_t width, uint32_t height, int quality);

/* ---- scene upload -------------------------------------------------------- */
/* The arrays are copied to the GPU asynchronously on the context's stream.  Pageable arrays are staged before the call
 * returns; PINNED arrays (srt_host_alloc, cudaHostAlloc) must stay valid and unmodified until srt_synchronize or the
 * next synchronous (host-pointer) call on the context returns. */
int  srt_set_meshes(srt_ctx* ctx, int n_meshes, const srt_mesh_desc* meshes);
int  srt_set_camera(srt_ctx* ctx, const srt_camera* cam);

/* ---- X-ray L-buffer path (the hot path) ---------------------------------- */
/*
 * Render rows [row_begin, row_end) of the L-buffer (pre-repair attenuated
 * intensity; -1 marks artefact pixels, lbuffer.cxx:805-809).  All output
 * pointers address FULL frames (width*height, row-major, index row*W+col as
 * lbuffer.cxx:254); only the band's rows are written.  path_len_out /
 * sign_sum_out / hit_count_out (mesh 0: summed signed path length `distance`,
 * `signSum`, number of accepted intersections) are optional (NULL).
 * Host-pointer version: synchronous, includes the device->host copy.
 */
int  srt_render_lbuffer(srt_ctx* ctx, const srt_params* params, uint32_t row_begin, uint32_t row_end,
                        float* L_out, float* path_len_out, int32_t* sign_sum_out, uint32_t* hit_count_out);
/* Device-pointer version: asynchronous on the context's stream (same layout).
 * Lets the caller run an in-place NCCL allgather over the bands. */
int  srt_render_lbuffer_device(srt_ctx* ctx, const srt_params* params, uint32_t row_begin, uint32_t row_end,
                               float* dL, float* d_path_len, int32_t* d_sign_sum, uint32_t* d_hit_count);

/*
 * Load-balanced multi-GPU sharding.  Contiguous bands are unbalanced (rows through the object run the exact
 * path far more often), so N ranks can instead take the 16-row blocks of the frame block-cyclically: rank
 * `part` renders blocks part, part + n_parts, ... into a DENSE slot buffer of srt_interleave_slot_rows() rows
 * (local row = (block / n_parts) * 16 + r).  After one all_gather of the slots ([n_parts][slot_rows][W]),
 * srt_deinterleave_device scatters them back to image order.  Results are bit-identical to a single render.
 */
int  srt_interleave_slot_rows(uint32_t height, uint32_t n_parts, uint32_t* slot_rows);
int  srt_render_lbuffer_interleaved_device(srt_ctx* ctx, const srt_params* params, uint32_t n_parts, uint32_t part,
                                           float* d_slot_L, float* d_slot_path_len, int32_t* d_slot_sign_sum,
                                           uint32_t* d_slot_hit_count);
int  srt_deinterleave_device(srt_ctx* ctx, uint32_t n_parts, const float* d_slots, float* d_frame);
/*
 * Fused render + assembly: rank `part` renders its 16-row blocks (same block-cyclic split) and the kernel stores
 * every L value at its IMAGE position (row * W + col) into each of the n_frames (1 ... 8) full-frame buffers in
 * d_frames -- the rank's own frame plus the peers' frames mapped over NVLink (CUDA IPC / symmetric memory), or a
 * single NVSwitch multicast mapping that replicates the store to every GPU (n_frames = 1 and
 * SRT_FLAG_FRAME0_MULTICAST in params->flags).  The transfer overlaps the render tile by tile; no all_gather and
 * no de-interleave pass follow, only a barrier before srt_repair_device.
 * Replaces the MPI_Send / MPI_Recv gather of src/main-mpi.cxx:854-881.
 */
int  srt_render_lbuffer_blocks_device(srt_ctx* ctx, const srt_params* params, uint32_t n_parts, uint32_t part,
                                      uint32_t n_frames, float* const* d_frames);

/*
 * Device-side barrier over the n_ranks GPUs of one node (one process per GPU), on the context's stream: d_flags[j] is
 * rank j's flag array (n_ranks 32-bit words in peer-mapped memory, zero before the first use) as mapped in THIS
 * process; epoch starts at 1 and grows by one per barrier on a given set of flag arrays.  The kernels launched before
 * it have stored their rows into the peers' frames; the kernels launched after it start when every rank's rows have
 * landed in this rank's frame.  A one-CTA kernel: any number of barriers may be in flight on different streams.
 *
 * srt_repair_blocks_device: the sharded repair pass -- the rank repairs the pixels of ITS 16-row blocks (reading the
 * assembled L frame d_L: the +-4-pixel stencil reaches into rows the neighbouring ranks rendered) and stores them into
 * each of the n_images image buffers (own first, then the peers' mappings; or one multicast mapping with
 * SRT_FLAG_FRAME0_MULTICAST), so that no rank spends time on pixels another rank repairs.
 *
 * srt_render_xray_blocks_device: one frame of the sharded pipeline in one call (the gather of
 * src/main-mpi.cxx:854-881 fused into the kernels; every rank ends up with the complete repaired image):
 *   srt_render_lbuffer_blocks_device -> barrier (epoch 2 e - 1) -> srt_repair_blocks_device(d_own_frame) -> barrier (2 e),
 * e = `epoch` = 1, 2, ... counting the frames issued on this set of flag arrays.  When the call's work has completed on
 * the stream, d_images (the rank's own one) holds the whole repaired image.
 */
int  srt_peer_barrier_device(srt_ctx* ctx, uint32_t n_ranks, uint32_t rank, uint32_t* const* d_flags, uint32_t epoch,
                             uint32_t phases /* 3 = signal + wait; 1 = signal only, 2 = wait only: a caller driving several
                                                ranks from ONE thread enqueues every rank's signal before the first wait */);
int  srt_repair_blocks_device(srt_ctx* ctx, const srt_params* params, uint32_t n_parts, uint32_t part, const float* d_L,
                              uint32_t n_images, float* const* d_images);
int  srt_render_xray_blocks_device(srt_ctx* ctx, const srt_params* params, uint32_t n_parts, uint32_t part,
                                   uint32_t n_frames, float* const* d_frames, uint32_t* const* d_flags, uint32_t epoch,
                                   const float* d_own_frame, uint32_t n_images, float* const* d_images);

/* Artefact repair + image fill over a full frame (lbuffer.cxx:329-404). L_in and image_out must not alias. */
int  srt_repair(srt_ctx* ctx, const srt_params* params, const float* L_in, float* image_out);
int  srt_repair_device(srt_ctx* ctx, const srt_params* params, const float* dL_in, float* d_image_out);

/* Whole frame in one call: L-buffer for all rows, repair, image to host
 * (what renderLoop + the output loop of main() produce; Image::getData layout). */
int  srt_render_xray(srt_ctx* ctx, const srt_params* params, float* image_out);

/* ---- sort-based X-ray (the serial driver, src/main.cxx:626-743; what produced out/dragon-128x128-serial.txt) ----
 * Per ray: the intersection distances t > 1e-7 with mesh 0 are sorted, exit - entry pairs summed (an odd count counts
 * as no path, :714-718), image = incident * exp(-(mu0 * (d * length_to_cm))) in the reference's float / double mix
 * (mu0 = meshes[0].mu: the driver hard-codes 0.3971f).  No artefact flag, no repair pass.  Runs on the filtered path
 * (srt_params.kernel AUTO / FILTERED / CULLED are all served by it). */
int  srt_render_xray_sorted(srt_ctx* ctx, const srt_params* params, float* image_out);
int  srt_render_xray_sorted_device(srt_ctx* ctx, const srt_params* params, float* d_image, float* d_path_len_or_null,
                                   uint32_t* d_hit_count_or_null);

/* ---- one process, N GPUs (replaces the MPI driver, src/main-mpi.cxx:258-293 partition + :854-881 gather) ----
 *
 * srt_multi owns one context per device.  A frame is rendered in block-cyclic 16-row blocks (device i renders blocks
 * i, i + N, ...); every render kernel stores its L rows into its own frame and the rows the neighbours' repair stencils
 * reach into the frames of the two neighbouring devices over NVLink peer mappings (cudaDeviceEnablePeerAccess);
 * cross-device events order the renders before the sharded repair passes (each device repairs its own blocks and stores
 * them into the image of EVERY device), and a second event barrier ends the frame: each device ends up with the
 * complete, repaired image -- bit-identical to a single-GPU render.  Scene and camera are replicated.  Requires peer
 * access between all the devices (an NVLink / NVSwitch box); 1 <= n_devices <= 8.
 */
typedef struct srt_multi srt_multi;
int  srt_multi_create(srt_multi** out, const int* device_ids, int n_devices);
void srt_multi_destroy(srt_multi* m);
const char* srt_multi_last_error(const srt_multi* m);     /* m may be NULL: last create error */
int  srt_multi_device_count(const srt_multi* m);
int  srt_multi_set_meshes(srt_multi* m, int n_meshes, const srt_mesh_desc* meshes);
int  srt_multi_set_camera(srt_multi* m, const srt_camera* cam);
/* Render + assemble + repair on all devices; asynchronous.  After srt_multi_synchronize the device pointers returned
 * by srt_multi_frame hold, for the LAST render, the complete repaired image (d_image, on every device) and the device's
 * part of the pre-repair L-buffer (d_L: its own blocks plus the neighbours' halo rows; complete only with one device). */
int  srt_multi_render_xray_device(srt_multi* m, const srt_params* params);
int  srt_multi_synchronize(srt_multi* m);
int  srt_multi_frame(srt_multi* m, int index, float** d_L, float** d_image);
/* The same, synchronous, image copied to host memory from the first device (what renderLoop + the gather leave on
 * rank 0). */
int  srt_multi_render_xray(srt_multi* m, const srt_params* params, float* image_out);
/* Device time of the last render per device: render kernels (index i) and the whole step of the slowest device. */
int  srt_multi_last_timing(srt_multi* m, float* render_ms_per_device, float* step_ms);

/* ---- shaded variant (src/main-omp.cxx:588-797; same loop in main-mpi*.cxx / main-pthreads-rows.cxx) ---- */
/*
 * Closest hit over ALL meshes (z-buffer, strict `>`, ascending triangle order, negative t accepted), Phong
 * shading with the hit triangle's cached normal and its mesh's material, shadow ray against every other
 * triangle (t > 1e-7 => colour * 0.25; "any blocker" does not depend on the order, so each ray only runs the
 * reference's test on the triangles whose projection from the light can cover it -- a grid in light space, see
 * kernels.h ShadowGrid), area-barycentric nearest-texel lookup for textured meshes, clamp +
 * int() truncation to 8 bits.  rgb_out: width*height*3 bytes, row-major; pixels without a hit get
 * background_rgb (the reference's Image(width, height, r, g, b) initial colour).  The camera's
 * pixel_spacing must be the one of the shaded drivers (from the bbox of all meshes, :594-607).
 */
int  srt_render_shaded(srt_ctx* ctx, const srt_params* params, const srt_light* light,
                       const uint8_t background_rgb[3], uint8_t* rgb_out);
int  srt_render_shaded_device(srt_ctx* ctx, const srt_params* params, const srt_light* light,
                              const uint8_t background_rgb[3], uint8_t* d_rgb, uint8_t* d_shadow_or_null);
/* Diagnostic: z-buffer, scene-wide triangle index and shadow flag (1 = shadow ray blocked) per pixel of the last shaded
 * render; any pointer may be NULL.  The flags are only kept when the render was given no shadow buffer of its own. */
int  srt_debug_shaded_buffers(srt_ctx* ctx, float* z_out, int32_t* tri_id_out, uint8_t* shadow_out);
/* Diagnostic: cells per side of the light-space grid the shadow rays look their blockers up in (0 = heuristic,
 * < 0 = no grid: every shadow ray tests every triangle like the reference's loop; identical images). */
int  srt_debug_force_shadow_grid(srt_ctx* ctx, int cells_per_side);

/* Last-call timings and counters (waits for the last render on the stream to finish). */
int  srt_last_timing(srt_ctx* ctx, srt_timing* out);
/* Diagnostic: make renders count the pairs that reach the exact path (costs one atomic per CTA). */
int  srt_debug_count_exact(srt_ctx* ctx, int enable);
/* Diagnostic: force the number of box columns a CTA of the filtered kernel covers (grid.x granularity); 0 = heuristic. */
int  srt_debug_force_scan_cols(srt_ctx* ctx, int cols);
/* Diagnostic: cap the per-bin triangle lists of the culled kernel and the two candidate lists of the filtered path
 * (entries; 0 = no cap).  Overflowing lists fall back to exact but slow paths; tests use this to exercise them. */
int  srt_debug_force_list_caps(srt_ctx* ctx, uint32_t bin_cap, uint32_t cand_cap, uint32_t pix_cap);
/* Diagnostic: measured FP32 FMA throughput of the context's GPU in TFLOP/s. */
int  srt_debug_ffma_peak(srt_ctx* ctx, int repeats, float* tflops_out);
/* Diagnostic: padding granularity of packed meshes (triangles) and the default box of the filtered kernel. */
int  srt_debug_pack_info(uint32_t* pad_to, uint32_t* box_w, uint32_t* box_h);
/* Diagnostic, host only (no GPU needed): the packed per-triangle constants the kernels stream.
 * exact16: 16 floats / triangle; filt12: 12, filt2_8: 8 floats / padded triangle (filt12[7], [11] = the box bounds R, R2
 * of the two level-0 strips for a box of box_w x box_h pixels, 0 = default box).  Any output may be NULL. */
int  srt_debug_prepack(const float* vertices, uint32_t n_triangles, const srt_camera* cam, uint32_t box_w, uint32_t box_h,
                       float* exact16, float* filt12, float* filt2_8, uint32_t* n_padded, uint32_t* n_unfilterable);
/* Diagnostic, host only (no GPU needed): the thread pool behind srt_set_host_copy_threads copies `bytes` from src to dst
 * in chunks of chunk_bytes with `threads` helper threads (0 = the calling thread alone) that poll for work for spin_us
 * microseconds before they sleep (0 = sleep at once). */
int  srt_debug_host_copy(void* dst, const void* src, size_t bytes, size_t chunk_bytes, int threads, int spin_us);
/* Diagnostic: bytes per DMA chunk of that path (0 = default 1 MiB; a power of two from 64 KiB to 32 MiB) and the time an
 * idle copy thread polls for the next chunk of a frame before it goes back to sleep (-1 = default 2000 us, 0 = never). */
int  srt_debug_tune_host_copy(srt_ctx* ctx, size_t chunk_bytes, int spin_us);
/* Diagnostic: number of device-to-host copies of the context that went through the pinned staging area so far. */
int  srt_debug_host_copy_stats(srt_ctx* ctx, uint64_t* staged_copies);
/* Diagnostic: the same arrays as computed ON THE GPU for mesh `mesh` of the context's scene. */
int  srt_debug_download_packed(srt_ctx* ctx, int mesh, float* exact16, float* filt12, float* filt2_8);

#ifdef __cplusplus
}
#endif
#endif /* SRT_H */
