// srt_api.cpp -- the extern "C" boundary declared in include/srt.h.
//
// Host responsibilities: validate arguments, keep the scene resident on the GPU (packed by
// prepack.cpp), compute the camera with the reference's exact float operation order, launch
// the kernels on the context's stream and time them with CUDA events.  No compute happens
// on the CPU: if CUDA is unavailable every render call fails with SRT_ERR_NO_DEVICE /
// SRT_ERR_CUDA.  Compile with -ffp-contract=off (camera arithmetic must match the
// reference bit for bit).
#include "srt.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <nvjpeg.h>

#include "host_copy.h"
#include "kernels.h"
#include "prepack.h"

#if defined(__FMA__) && !defined(SRT_ALLOW_FMA_HOST)
#error "build srt_api.cpp without -mfma/-march=native"
#endif

#ifndef SRT_RESIDENT_CTAS
#define SRT_RESIDENT_CTAS 4      // CTAs of the filtered kernel resident per SM if the occupancy query fails
#endif
#ifndef SRT_BOX_PX
#define SRT_BOX_PX 64            // default box of the filtered kernel: SRT_BOX_PX x SRT_BOX_PY pixels, SRT_BOX_NB boxes per thread
#endif
#ifndef SRT_BOX_PY
#define SRT_BOX_PY 16
#endif
#ifndef SRT_SCAN_K
#define SRT_SCAN_K 4             // triangles a thread of the filtered kernel holds in registers
#endif


namespace {

constexpr uint32_t kMaxScanK = 8;
constexpr size_t kScanSlack = 128 * kMaxScanK;   // the chunk-major scan array is padded to whole chunks

thread_local std::string g_create_error;

// One mesh resident on the device: raw geometry as uploaded + the packed arrays the kernels stream.
struct MeshBuf {
    float* d_vertices = nullptr;     // 9 per triangle
    float* d_normals = nullptr;      // 3 per triangle, only when the caller supplied normals
    void* filt = nullptr; void* exact = nullptr;   // filt: cap level-0 records (48 B) followed by cap level-1 records (32 B)
    void* scan = nullptr;            // chunk-major copy of the level-0 records for the filtered kernel (cap + kScanSlack slots)
    uint32_t n_chunks = 0;           // chunks of 128 * scan_k triangles of the current packing
    const float4* filt0() const { return (const float4*)filt; }
    const float4* filt1() const { return (const float4*)((const char*)filt + cap * sizeof(srt::FilterTri)); }
    size_t cap = 0;                  // capacity in triangles of all of the above
    uint32_t n = 0, n_padded = 0;
    bool has_normals = false;
    float mu = 0.0f;
    // shaded variant
    float* d_texcoords = nullptr; size_t texcoord_cap = 0; bool has_texcoords = false;
    uint8_t* d_tex = nullptr; size_t tex_cap = 0; uint32_t tex_w = 0, tex_h = 0;
    srt_material material{};
};

} // namespace

struct srt_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };   // render begin / end, repair (shade) begin / end, main kernel begin / end, shadow grid begin
    std::string error;

    std::vector<MeshBuf> meshes;
    size_t n_meshes = 0;                 // meshes in use (buffers beyond are kept for reuse)
    bool have_camera = false;
    srt_camera camera{};
    bool packed = false;                 // packed arrays match meshes + camera + box shape
    uint32_t box_px = 0, box_py = 0;     // box of the filtered kernel the packed level-0 bounds (FilterTri::R) are for
    uint32_t scan_k = 0;                 // triangles per thread the chunk-major scan arrays are laid out for
    bool filterable = false;

    std::vector<srt::ShadeMesh> h_shade; // shaded variant: per-mesh descriptors
    srt::ShadeMesh* d_shade_meshes = nullptr; size_t shade_cap = 0;
    uint8_t* d_rgb = nullptr; size_t rgb_cap = 0;
    // shadow rays: light-space grid of the scene (kernels.h ShadowGrid), rebuilt when the scene or the light changes
    uint8_t* d_shadow = nullptr; size_t shadow_cap = 0;
    srt::ShadowGrid grid{};
    uint64_t scene_serial = 1, grid_scene = 0, bbox_scene = 0;
    float grid_light[3] = { 0, 0, 0 }, scene_bbox[6] = { 0, 0, 0, 0, 0, 0 };
    uint32_t* d_bbox = nullptr;
    size_t grid_cells_cap = 0, grid_tris_cap = 0, grid_global_cap = 0;
    int force_shadow_grid = 0;           // diagnostic: 0 = heuristic, > 0 = cells per side, < 0 = brute force
    std::vector<srt::DevMesh> h_desc;    // host copy of the descriptor arrays (default | all contribute)
    srt::DevMesh* d_meshes = nullptr;    // device array: [0, n) default semantics, [n, 2n) all meshes contribute
    size_t desc_cap = 0;
    float* d_u_off = nullptr; float* d_v_off = nullptr;
    size_t off_cap_u = 0, off_cap_v = 0;
    uint64_t total_tris = 0, mesh0_tris = 0;

    // frame scratch for the host-pointer entry points
    size_t frame_cap = 0;
    float* d_L = nullptr; float* d_img = nullptr; float* d_dist = nullptr; int32_t* d_ssum = nullptr; uint32_t* d_hits = nullptr;
    unsigned long long* d_counters = nullptr;
    bool count_exact = false;
    // per-pixel hit lists of the filtered kernel (counts are zero between renders: resolve_hits_kernel resets them)
    uint32_t* d_hit_head = nullptr; uint32_t* d_hit_next = nullptr; size_t hit_pixels_cap = 0;
    // candidate list of the filtered kernel: survivors of the scan.  h_cand_count (pinned) receives the count of every
    // render asynchronously; a later render grows the list when the last known count came close to the capacity.
    uint2* d_cand = nullptr; uint2* d_pix = nullptr; uint32_t* d_cand_count = nullptr; uint32_t cand_cap = 0, pix_cap = 0;
    uint2* d_big = nullptr; uint32_t big_cap = 0;       // big-box shapes: survivors of the scan, refined into d_cand
    uint32_t* h_cand_count = nullptr;    // [0] box list, [1] pixel list, [2] big-box list
    uint64_t lists_key = 0;              // launch geometry whose list sizes have been verified (0 = none)
    uint32_t verified_cand = 0, verified_pix = 0, verified_big = 0;   // ... and those sizes (they never change for a scene / camera / geometry)
    int sms = 148;                       // multiprocessors of the device (queried once)
    uint64_t resident_key = ~0ull; int resident_val = 0;   // cached occupancy of the scan kernel for (shape, residency cap)
    int cnt_parity = 0;                  // which pair of the double-buffered list counters (d_cand_count[4]) is clean
    uint32_t scan_max_ctas = 0;          // cap on the scan's resident CTAs per SM (0 = none; srt_set_scan_residency)
    bool profiling = true;               // record the per-render timing events (srt_set_profiling)
    // one frame of the sharded pipeline (srt_render_xray_blocks_device) as a CUDA graph: the call's launches are captured the
    // second time a call signature comes round in the steady state and replayed from then on (a rank's share of an 8-GPU frame
    // is ~40 us of device work behind ~8 launches: the host's launch cost was the bound)
    struct FrameGraph { std::vector<uint64_t> sig; uint32_t steady_runs = 0; uint64_t lists_key_after = 0; bool flips_parity = false; cudaGraphExec_t exec = nullptr; };
    std::vector<FrameGraph> frame_graphs;
    uint64_t pack_serial = 0;            // bumped whenever the packed arrays are rebuilt (new scene / camera / box shape)
    bool verified_last_render = false;   // the last render_device call waited for list counts (first render of a launch geometry)
    uint32_t* d_barrier_ctr = nullptr;   // epoch counter of the fused call's barriers (device memory), mirrored in barrier_epochs
    uint64_t barrier_epochs = 0;
    int force_scan_cols = 0;             // diagnostic override of the box columns per CTA: 0 = heuristic
    uint32_t force_bin_cap = 0, force_cand_cap = 0, force_pix_cap = 0;    // diagnostic: shrink the lists to exercise the overflow paths
    // per-bin triangle lists of the culled kernel
    uint32_t* d_bin_entries = nullptr; uint32_t* d_bin_counts = nullptr; size_t bin_entries_cap = 0, bin_counts_cap = 0;
    bool counters_pending = false, timing_pending = false;
    // frames returned into PAGEABLE host buffers (copy_to_host): a pinned staging area the DMA engine fills chunk by chunk,
    // one event per chunk, and a few threads that move the chunks that have landed into the caller's buffer (host_copy.h)
    void* h_stage = nullptr; size_t stage_cap = 0;
    std::vector<cudaEvent_t> stage_ev;
    std::unique_ptr<srt::HostCopyPool> copy_pool;
    int copy_threads = -1;               // threads besides the caller's: -1 = from the core count, 0 = leave pageable copies to the driver
    size_t copy_chunk = 0;               // bytes per DMA chunk; 0 = kStageChunk (srt_debug_tune_host_copy)
    int copy_spin_us = -1;               // how long an idle copy thread polls for the next chunk; -1 = kCopySpinUs
    uint64_t staged_copies = 0;          // copies that went through the staging area (srt_debug_host_copy_stats)

    srt_timing timing{};
};

namespace {

int fail(srt_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->error = buf; else g_create_error = buf;
    return code;
}

#define SRT_CUDA(c, call)                                                                        \
    do { cudaError_t e__ = (call);                                                               \
         if (e__ != cudaSuccess) return fail((c), SRT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); } while (0)

// elapsed time between two of the context's events; 0 when the per-render events are switched off (srt_set_profiling)
int elapsed_ms(srt_ctx* c, float* out, int from, int to)
{
    *out = 0.0f;
    if (!c->profiling) return SRT_OK;
    SRT_CUDA(c, cudaEventElapsedTime(out, c->ev[from], c->ev[to]));
    return SRT_OK;
}

void free_mesh(MeshBuf& m)
{
    cudaFree(m.d_vertices); cudaFree(m.d_normals); cudaFree(m.filt); cudaFree(m.exact); cudaFree(m.scan);
    float* tc = m.d_texcoords; size_t tcc = m.texcoord_cap; uint8_t* tx = m.d_tex; size_t txc = m.tex_cap;
    m = MeshBuf{};
    m.d_texcoords = tc; m.texcoord_cap = tcc; m.d_tex = tx; m.tex_cap = txc;       // kept: sized independently
}

void free_scene(srt_ctx* c)
{
    for (auto& m : c->meshes) { free_mesh(m); cudaFree(m.d_texcoords); cudaFree(m.d_tex); }
    cudaFree(c->d_shade_meshes); c->d_shade_meshes = nullptr; c->shade_cap = 0;
    cudaFree(c->d_rgb); c->d_rgb = nullptr; c->rgb_cap = 0;
    cudaFree(c->d_shadow); c->d_shadow = nullptr; c->shadow_cap = 0;
    cudaFree(c->grid.cell_off); cudaFree(c->grid.cell_cursor); cudaFree(c->grid.cell_tris); cudaFree(c->grid.global_tris);
    cudaFree(c->grid.n_global); cudaFree(c->d_bbox);
    c->grid = srt::ShadowGrid{}; c->d_bbox = nullptr; c->grid_cells_cap = c->grid_tris_cap = c->grid_global_cap = 0;
    c->grid_scene = c->bbox_scene = 0;
    c->meshes.clear(); c->n_meshes = 0;
    cudaFree(c->d_meshes); c->d_meshes = nullptr; c->desc_cap = 0;
    cudaFree(c->d_u_off); c->d_u_off = nullptr; c->off_cap_u = 0;
    cudaFree(c->d_v_off); c->d_v_off = nullptr; c->off_cap_v = 0;
    c->packed = false;
}

void free_frame(srt_ctx* c)
{
    cudaFree(c->d_L); cudaFree(c->d_img); cudaFree(c->d_dist); cudaFree(c->d_ssum); cudaFree(c->d_hits);
    c->d_L = c->d_img = c->d_dist = nullptr; c->d_ssum = nullptr; c->d_hits = nullptr; c->frame_cap = 0;
    cudaFree(c->d_hit_head); cudaFree(c->d_hit_next);
    c->d_hit_head = nullptr; c->d_hit_next = nullptr; c->hit_pixels_cap = 0;
    cudaFree(c->d_cand); cudaFree(c->d_pix); cudaFree(c->d_big); cudaFree(c->d_cand_count); if (c->h_cand_count) cudaFreeHost(c->h_cand_count);
    c->d_cand = nullptr; c->d_pix = nullptr; c->d_big = nullptr; c->d_cand_count = nullptr; c->h_cand_count = nullptr;
    c->cand_cap = 0; c->pix_cap = 0; c->big_cap = 0;
    cudaFree(c->d_bin_entries); cudaFree(c->d_bin_counts);
    c->d_bin_entries = nullptr; c->d_bin_counts = nullptr; c->bin_entries_cap = 0; c->bin_counts_cap = 0;
}

int ensure_frame(srt_ctx* c, size_t pixels)
{
    if (pixels <= c->frame_cap) return SRT_OK;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    free_frame(c);
    SRT_CUDA(c, cudaMalloc(&c->d_L, pixels * 4));
    SRT_CUDA(c, cudaMalloc(&c->d_img, pixels * 4));
    SRT_CUDA(c, cudaMalloc(&c->d_dist, pixels * 4));
    SRT_CUDA(c, cudaMalloc(&c->d_ssum, pixels * 4));
    SRT_CUDA(c, cudaMalloc(&c->d_hits, pixels * 4));
    c->frame_cap = pixels;
    return SRT_OK;
}

// ---- device -> caller's host buffer ---------------------------------------------------------------------------------
constexpr size_t kStageChunk = 1u << 20;        // bytes per DMA chunk (one event each)
constexpr size_t kStageBytes = 32u << 20;       // pinned staging area of a context; larger frames go through it in rounds
constexpr size_t kStageMin = 2u << 20;          // smaller copies are left to the driver
constexpr int kCopySpinUs = 2000;               // copy threads poll this long for a frame's chunks before they sleep again

int default_copy_threads()
{
    static const int n = [] {
        if (const char* e = std::getenv("SRT_COPY_THREADS")) return std::max(0, std::min(32, std::atoi(e)));
        // measured on the B200 boxes (16 cores): 16 MiB move at 17 / 32 / 44 / 56 / 59 / 43 GB/s with 0 / 1 / 2 / 4 / 7 / 12
        // helpers, and the whole srt_render_xray call is shortest with 4 (gpurun_out/r02e_*): a quarter of the cores, at most 4
        const unsigned hw = std::thread::hardware_concurrency();      // 0 = unknown
        return (int)std::min(4u, hw / 4);
    }();
    return n;
}

void free_stage(srt_ctx* c)
{
    c->copy_pool.reset();
    for (auto& ev : c->stage_ev) if (ev) cudaEventDestroy(ev);
    c->stage_ev.clear();
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr; c->stage_cap = 0;
}

// Staging area, events and threads for `threads` helpers; false = not available (the caller falls back to the driver's copy).
bool ensure_stage(srt_ctx* c, int threads)
{
    if (!c->h_stage) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, kStageBytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
        c->h_stage = p; c->stage_cap = kStageBytes;
    }
    const size_t chunk = c->copy_chunk ? c->copy_chunk : kStageChunk;
    const size_t n_ev = (c->stage_cap + chunk - 1) / chunk;
    while (c->stage_ev.size() < n_ev) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        c->stage_ev.push_back(ev);
    }
    const unsigned spin = (unsigned)(c->copy_spin_us >= 0 ? c->copy_spin_us : kCopySpinUs);
    if (!c->copy_pool || c->copy_pool->workers() != threads || c->copy_pool->spin_us() != spin) {
        c->copy_pool.reset();
        c->copy_pool.reset(new (std::nothrow) srt::HostCopyPool(threads, spin));
        if (!c->copy_pool) return false;
    }
    return true;
}

// Copy `bytes` from device memory to the caller's host buffer behind the work already on the context's stream.
//   * page-locked destination (srt_host_alloc / cudaHostAlloc / registered), small copies, copy threads switched off:
//     one cudaMemcpyAsync; the caller synchronises the stream afterwards, as before;
//   * pageable destination of kStageMin bytes or more: DMA chunks into the pinned staging area, and the copy threads move
//     every chunk that has landed while the next ones are in flight.  Returns with the bytes in place.
// The bytes that arrive are the same either way (tests/test_gpu_parity.py compares the two).
int copy_to_host(srt_ctx* c, void* dst, const void* d_src, size_t bytes)
{
    if (!bytes) return SRT_OK;
    const int threads = c->copy_threads >= 0 ? c->copy_threads : default_copy_threads();
    bool staged = false;
    if (bytes >= kStageMin && threads > 0) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, dst) == cudaSuccess) staged = at.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (staged && !ensure_stage(c, threads)) staged = false;
    if (!staged) {
        SRT_CUDA(c, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
        return SRT_OK;
    }
    const char* src = (const char*)d_src;
    char* out = (char*)dst;
    char* stage = (char*)c->h_stage;
    const size_t chunk = c->copy_chunk ? c->copy_chunk : kStageChunk;
    cudaError_t err = cudaSuccess;
    for (size_t done = 0; done < bytes && err == cudaSuccess; ) {
        const size_t round = std::min(bytes - done, c->stage_cap);
        const size_t n_chunks = (round + chunk - 1) / chunk;
        size_t issued = 0;
        c->copy_pool->activate();                           // the threads wake up while the stream is still busy
        for (; issued < n_chunks && err == cudaSuccess; ++issued) {
            const size_t off = issued * chunk, n = std::min(chunk, round - off);
            err = cudaMemcpyAsync(stage + off, src + done + off, n, cudaMemcpyDeviceToHost, c->stream);
            if (err == cudaSuccess) err = cudaEventRecord(c->stage_ev[issued], c->stream);
            if (err != cudaSuccess) break;
        }
        for (size_t i = 0; i < issued && err == cudaSuccess; ++i) {
            const size_t off = i * chunk, n = std::min(chunk, round - off);
            err = cudaEventSynchronize(c->stage_ev[i]);
            if (err == cudaSuccess) c->copy_pool->submit(out + done + off, stage + off, n);
        }
        c->copy_pool->drain();                              // also on errors: no copy may outlive this call
        done += round;
    }
    if (err != cudaSuccess) return fail(c, SRT_ERR_CUDA, "staged device-to-host copy: %s", cudaGetErrorString(err));
    ++c->staged_copies;
    return SRT_OK;
}

// Grow a mesh's device buffers to hold n triangles (only reallocates when growing).
int ensure_mesh(srt_ctx* c, MeshBuf& m, size_t n, bool want_normals)
{
    const size_t pad = (n + srt::kPadTris - 1) / srt::kPadTris * srt::kPadTris;
    if (pad > m.cap || m.cap == 0) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        const bool had_normals = m.d_normals != nullptr;
        free_mesh(m);
        const size_t cap = pad + pad / 8 + 64;
        SRT_CUDA(c, cudaMalloc(&m.d_vertices, cap * 9 * sizeof(float)));
        SRT_CUDA(c, cudaMalloc(&m.filt, cap * (sizeof(srt::FilterTri) + sizeof(srt::FilterTri2))));
        SRT_CUDA(c, cudaMalloc(&m.exact, cap * sizeof(srt::ExactTri)));
        SRT_CUDA(c, cudaMalloc(&m.scan, (cap + kScanSlack) * sizeof(srt::FilterTri)));
        if (had_normals || want_normals) SRT_CUDA(c, cudaMalloc(&m.d_normals, cap * 3 * sizeof(float)));
        m.cap = cap;
    }
    if (want_normals && !m.d_normals) SRT_CUDA(c, cudaMalloc(&m.d_normals, m.cap * 3 * sizeof(float)));
    return SRT_OK;
}

// Work lists of the filtered path: at least `cand` {box, triangle} and `pix` {pixel, triangle} entries (only ever grows).
int ensure_lists(srt_ctx* c, uint64_t cand, uint64_t pix, uint64_t big = 0)
{
    cand = std::min<uint64_t>(cand, 0xffffffffull); pix = std::min<uint64_t>(pix, 0xffffffffull); big = std::min<uint64_t>(big, 0xffffffffull);
    if (!c->h_cand_count) {
        SRT_CUDA(c, cudaHostAlloc(&c->h_cand_count, 4 * sizeof(uint32_t), cudaHostAllocDefault));
        c->h_cand_count[0] = c->h_cand_count[1] = c->h_cand_count[2] = c->h_cand_count[3] = 0;
    }
    if (!c->d_cand_count) {                                    // two sets of {cand, pix, big, -}
        SRT_CUDA(c, cudaMalloc(&c->d_cand_count, 8 * sizeof(uint32_t)));
        SRT_CUDA(c, cudaMemsetAsync(c->d_cand_count, 0, 8 * sizeof(uint32_t), c->stream));
        c->cnt_parity = 0;
    }
    if (big > c->big_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_big); c->d_big = nullptr; c->big_cap = 0;
        SRT_CUDA(c, cudaMalloc(&c->d_big, big * sizeof(uint2)));
        c->big_cap = (uint32_t)big;
    }
    if (cand > c->cand_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_cand); c->d_cand = nullptr; c->cand_cap = 0;
        SRT_CUDA(c, cudaMalloc(&c->d_cand, cand * sizeof(uint2)));
        c->cand_cap = (uint32_t)cand;
    }
    if (pix > c->pix_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_pix); cudaFree(c->d_hit_next); c->d_pix = nullptr; c->d_hit_next = nullptr; c->pix_cap = 0;
        SRT_CUDA(c, cudaMalloc(&c->d_pix, pix * sizeof(uint2)));
        SRT_CUDA(c, cudaMalloc(&c->d_hit_next, pix * sizeof(uint32_t)));
        c->pix_cap = (uint32_t)pix;
    }
    return SRT_OK;
}

// (Re)build the packed arrays for the current meshes + camera + shape of the filtered kernel: all on the GPU,
// all asynchronous.
int pack_scene(srt_ctx* c, uint32_t box_px = SRT_BOX_PX, uint32_t box_py = SRT_BOX_PY, uint32_t scan_k = SRT_SCAN_K)
{
    if (c->packed && c->box_px == box_px && c->box_py == box_py && c->scan_k == scan_k) return SRT_OK;
    if (scan_k == 0 || scan_k > kMaxScanK) return fail(c, SRT_ERR_INVALID, "triangles per thread %u out of range", scan_k);
    if (c->n_meshes == 0) return fail(c, SRT_ERR_INVALID, "srt_set_meshes has not been called");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    const srt_camera& cam = c->camera;
    const uint32_t W = cam.width, H = cam.height;
    if (W > c->off_cap_u) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_u_off); c->d_u_off = nullptr;
        SRT_CUDA(c, cudaMalloc(&c->d_u_off, (size_t)W * 4)); c->off_cap_u = W;
    }
    if (H > c->off_cap_v) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_v_off); c->d_v_off = nullptr;
        SRT_CUDA(c, cudaMalloc(&c->d_v_off, (size_t)H * 4)); c->off_cap_v = H;
    }
    SRT_CUDA(c, srt::launch_offsets(c->d_u_off, c->d_v_off, W, H, cam.pixel_spacing[0], cam.pixel_spacing[1], c->stream));

    const srt::FrameGeom g = srt::make_frame_geom(cam.origin, cam.detector, cam.up, cam.right,
                                                  cam.pixel_spacing[0], cam.pixel_spacing[1], W, H, box_px, box_py);
    c->filterable = g.filterable != 0;

    const size_t nm = c->n_meshes;
    c->h_desc.assign(2 * nm, srt::DevMesh{});
    c->h_shade.assign(nm, srt::ShadeMesh{});
    c->total_tris = 0;
    for (size_t m = 0; m < nm; ++m) {
        MeshBuf& mb = c->meshes[m];
        srt::ShadeMesh& sm = c->h_shade[m];
        sm.vertices = mb.d_vertices; sm.exact = (const float4*)mb.exact;
        sm.texcoords = mb.has_texcoords ? mb.d_texcoords : nullptr;
        sm.tex = (mb.tex_w && mb.tex_h) ? mb.d_tex : nullptr; sm.tex_w = mb.tex_w; sm.tex_h = mb.tex_h;
        sm.n = mb.n; sm.first = (uint32_t)c->total_tris;
        for (int k = 0; k < 3; ++k) { sm.ka[k] = mb.material.ambient[k]; sm.kd[k] = mb.material.diffuse[k]; sm.ks[k] = mb.material.specular[k]; }
        sm.shininess = mb.material.shininess;
        mb.n_chunks = (mb.n_padded + 128 * scan_k - 1) / (128 * scan_k);
        SRT_CUDA(c, srt::launch_prepack(mb.d_vertices, mb.has_normals ? mb.d_normals : nullptr, mb.n, mb.n_padded, mb.n_chunks,
                                        scan_k, cam.origin, g, mb.exact, (void*)mb.filt0(), (void*)mb.filt1(), mb.scan, c->stream));
        srt::DevMesh& d = c->h_desc[m];
        d.filt0 = mb.filt0(); d.filt1 = mb.filt1(); d.exact = (const float4*)mb.exact;
        d.scan = (const float4*)mb.scan; d.n_chunks = mb.n_chunks;
        d.n = mb.n; d.n_padded = mb.n_padded; d.mu = mb.mu;
        d.first = (uint32_t)c->total_tris;
        d.contributes = (m == 0);                           // lbuffer.cxx:788: only meshes[0] adds path length
        c->h_desc[nm + m] = d;
        c->h_desc[nm + m].contributes = 1;
        c->total_tris += mb.n;
    }
    c->mesh0_tris = c->meshes[0].n;
    if (2 * nm > c->desc_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_meshes); c->d_meshes = nullptr;
        SRT_CUDA(c, cudaMalloc(&c->d_meshes, 2 * nm * sizeof(srt::DevMesh))); c->desc_cap = 2 * nm;
    }
    SRT_CUDA(c, cudaMemcpyAsync(c->d_meshes, c->h_desc.data(), 2 * nm * sizeof(srt::DevMesh), cudaMemcpyHostToDevice, c->stream));
    if (nm > c->shade_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_shade_meshes); c->d_shade_meshes = nullptr;
        SRT_CUDA(c, cudaMalloc(&c->d_shade_meshes, nm * sizeof(srt::ShadeMesh))); c->shade_cap = nm;
    }
    SRT_CUDA(c, cudaMemcpyAsync(c->d_shade_meshes, c->h_shade.data(), nm * sizeof(srt::ShadeMesh), cudaMemcpyHostToDevice, c->stream));
    c->packed = true; c->box_px = box_px; c->box_py = box_py; c->scan_k = scan_k; ++c->pack_serial;
    c->lists_key = 0;                                       // new scene / camera: the work-list sizes are unverified again
    return SRT_OK;
}

// Shape of the filtered kernel: srt_params::pixels_per_thread = box width | box height << 8 | triangles per thread << 16
// (0 in a field = default).
bool decode_box(uint32_t v, uint32_t& px, uint32_t& py, uint32_t& k)
{
    px = v & 0xffu; py = (v >> 8) & 0xffu; k = (v >> 16) & 0xffu;
    if (px == 0) { px = SRT_BOX_PX; if (py == 0) py = SRT_BOX_PY; }
    if (py == 0) py = 1;
    if (k == 0) k = SRT_SCAN_K;
    return srt::filtered_shape_supported(px, py, k);
}

// parts == 0: contiguous band [row_begin, row_end) written at image positions.
// parts >= 1: block-cyclic part `part` of `parts` over the whole frame, written compactly (slot layout).
// extra_frames != nullptr: fused assembly, L goes to the image positions of dL and of the n_extra extra frames;
// multicast: dL is an NVSwitch multicast mapping (multimem.st).
int render_device(srt_ctx* c, const srt_params* params, uint32_t row_begin, uint32_t row_end,
                  float* dL, float* d_dist, int32_t* d_ssum, uint32_t* d_hits, uint32_t parts = 0, uint32_t part = 0,
                  int mode = 0, float* const* extra_frames = nullptr, uint32_t n_extra = 0, bool image_positions = false,
                  bool multicast = false, bool halo = false)
{
    if (!params) return fail(c, SRT_ERR_INVALID, "params is NULL");
    if (mode != 1 && !dL) return fail(c, SRT_ERR_INVALID, "L buffer is NULL");
    if (mode == 1 && (!d_dist || !d_ssum)) return fail(c, SRT_ERR_INVALID, "closest-hit pass needs z and id buffers");
    if (n_extra > 7 || (n_extra && !extra_frames)) return fail(c, SRT_ERR_INVALID, "at most 7 extra frames");
    uint32_t kernel = params->kernel;
    const int audit = (kernel == SRT_KERNEL_AUDIT);
    if (audit) kernel = SRT_KERNEL_EXACT;
    if (kernel != SRT_KERNEL_AUTO && kernel != SRT_KERNEL_EXACT && kernel != SRT_KERNEL_FILTERED && kernel != SRT_KERNEL_CULLED)
        return fail(c, SRT_ERR_INVALID, "unknown kernel selector %u", params->kernel);
    uint32_t box_px, box_py, scan_k;
    if (!decode_box(params->pixels_per_thread, box_px, box_py, scan_k))
        return fail(c, SRT_ERR_INVALID, "shape %u x %u pixels, %u triangles per thread (pixels_per_thread = 0x%x) is not built", box_px,
                    box_py, scan_k, params->pixels_per_thread);
    const bool repack = !(c->packed && c->box_px == box_px && c->box_py == box_py && c->scan_k == scan_k);
    int rc = pack_scene(c, box_px, box_py, scan_k);
    if (rc) return rc;
    const srt_camera& cam = c->camera;
    if (row_begin > row_end || row_end > cam.height)
        return fail(c, SRT_ERR_INVALID, "row band [%u, %u) outside the %u-row frame", row_begin, row_end, cam.height);

    srt::RenderArgs a{};
    a.cam.ox = cam.origin[0]; a.cam.oy = cam.origin[1]; a.cam.oz = cam.origin[2];
    a.cam.dx = cam.detector[0]; a.cam.dy = cam.detector[1]; a.cam.dz = cam.detector[2];
    a.cam.ux = cam.up[0]; a.cam.uy = cam.up[1]; a.cam.uz = cam.up[2];
    a.cam.rx = cam.right[0]; a.cam.ry = cam.right[1]; a.cam.rz = cam.right[2];
    a.cam.u_off = c->d_u_off; a.cam.v_off = c->d_v_off;
    a.cam.xc = (float)(0.5 - (double)cam.width / 2.0); a.cam.yc = (float)(0.5 - (double)cam.height / 2.0);
    a.cam.W = cam.width; a.cam.H = cam.height;
    const bool all = (params->flags & SRT_FLAG_ALL_MESHES_CONTRIBUTE) != 0 || mode == 1;
    a.mode = mode;
    a.meshes = c->d_meshes + (all ? c->n_meshes : 0);
    a.n_meshes = (int)c->n_meshes;
    a.row_begin = row_begin; a.row_end = row_end;
    const uint32_t total_blocks = (row_end - row_begin + 15) / 16;
    uint32_t band_rows = row_end - row_begin;
    if (parts == 0) { a.blk_stride = 1; a.blk_phase = 0; a.n_blocks = total_blocks; a.compact = 0; }
    else {
        a.blk_stride = parts; a.blk_phase = part; a.compact = 1;
        a.n_blocks = total_blocks > part ? (total_blocks - part + parts - 1) / parts : 0;
        band_rows = 0;
        for (uint32_t b = part; b < total_blocks; b += parts) band_rows += std::min(16u, row_end - row_begin - b * 16);
    }
    a.incident = params->incident; a.length_to_cm = params->length_to_cm;
    a.L = dL; a.dist = d_dist; a.ssum = d_ssum; a.hits = d_hits;
    a.L_image = image_positions ? 1 : 0;
    a.L_multicast = multicast ? 1 : 0;
    a.n_L_extra = n_extra;
    bool aligned = ((uintptr_t)dL & 15u) == 0 && (cam.width & 3u) == 0;
    for (uint32_t i = 0; i < n_extra; ++i) {
        if (!extra_frames[i]) return fail(c, SRT_ERR_INVALID, "extra frame %u is NULL", i);
        a.L_extra[i] = extra_frames[i];
        aligned = aligned && ((uintptr_t)extra_frames[i] & 15u) == 0;
    }
    a.L_vec4 = aligned ? 1 : 0;                         // 128-bit stores only into 16-byte aligned frames
    if (halo && parts > 1 && n_extra == parts - 1 && params->repair_depth >= 1 && params->repair_depth <= 16) {
        // extra_frames lists the peers in rank order (own rank skipped): keep the owners of the blocks above / below only
        const uint32_t prev = (part + parts - 1) % parts, next = (part + 1) % parts;
        a.L_extra[0] = extra_frames[prev < part ? prev : prev - 1];
        a.L_extra[1] = extra_frames[next < part ? next : next - 1];
        a.n_L_extra = 2; a.L_halo = (int)params->repair_depth;
    }
    a.box_px = box_px; a.box_py = box_py; a.scan_k = scan_k;
    a.counters = nullptr;
    if (c->count_exact || audit) {
        if (!c->d_counters) SRT_CUDA(c, cudaMalloc(&c->d_counters, 3 * sizeof(unsigned long long)));
        SRT_CUDA(c, cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(unsigned long long), c->stream));
        SRT_CUDA(c, cudaMemsetAsync(c->d_counters + 2, 0x7f, sizeof(unsigned long long), c->stream));      // two large floats (3.4e38)
        a.counters = c->d_counters;
    }

    // AUTO: the fastest kernel for the scene -- same results whichever runs.  A handful of triangles (config C4: the
    // 2-triangle logo quad) is fastest in the one-launch exact kernel (the multi-kernel paths cost more in launches than
    // such a frame in arithmetic); up to ~2e5 triangles the brute-force scan with 64 x 16 boxes wins (dragon, 2048^2:
    // 0.17 ms against 0.22 ms culled); beyond, the bins of the culled kernel pay off (998 000-triangle sphere at 8192^2:
    // 44 ms against 166 ms: every box still meets every triangle in the brute-force scan, and the bumpy silhouette's
    // long hit lists cost the resolve pass more than the culled kernel's in-order accumulation).
    if (kernel == SRT_KERNEL_AUTO)
        kernel = (c->total_tris <= 64 && mode == 0) ? SRT_KERNEL_EXACT : (c->total_tris <= 200000 ? SRT_KERNEL_FILTERED : SRT_KERNEL_CULLED);
    if (!c->filterable) kernel = SRT_KERNEL_EXACT;                        // camera outside the filter's preconditions
    if (mode == 2) {
        // the sort mode lives on the hit lists of the filtered path (an unfilterable camera lets every box through:
        // the lists overflow and the resolve pass orders each ray's intersections itself -- slow, still exact)
        if (params->kernel == SRT_KERNEL_EXACT || audit) return fail(c, SRT_ERR_INVALID, "the sort-based X-ray runs on SRT_KERNEL_FILTERED");
        kernel = SRT_KERNEL_FILTERED;
    }
    if (kernel == SRT_KERNEL_CULLED) {
        // Bins: 256 columns x 8 row blocks (128 rows) of a contiguous band; with block-cyclic rows the blocks of a rank
        // are `parts` blocks apart, so a bin is one block tall.  Lists: up to one slot per scene triangle per bin, capped
        // so that all lists together stay within 1 GiB (the 998 000-triangle sphere at 8192^2 has 2 048 bins: 131 072
        // slots each, ~15 000 used); a bin whose list overflows its share is rendered from the whole mesh instead.
        a.bin_blocks = a.blk_stride > 1 ? 1u : 8u;
        a.bins_x = (cam.width + srt::kBinCols - 1) / srt::kBinCols;
        a.bins_y = (a.n_blocks + a.bin_blocks - 1) / a.bin_blocks;
        const size_t bins = (size_t)a.bins_x * a.bins_y;
        if (c->total_tris >= (1ull << 32) || c->total_tris == 0) kernel = SRT_KERNEL_FILTERED;
        else if (bins > 0) {
            const uint64_t budget = (1ull << 30) / 4;                                      // entries
            const uint64_t per_bin = std::max<uint64_t>(std::min<uint64_t>(c->total_tris, budget / bins), std::min<uint64_t>(c->total_tris, 4096));
            a.bin_cap = (uint32_t)per_bin; a.total_tris = (uint32_t)c->total_tris;
            if (c->force_bin_cap) a.bin_cap = std::min<uint32_t>(a.bin_cap, c->force_bin_cap);
            if (bins * per_bin > c->bin_entries_cap) {
                SRT_CUDA(c, cudaStreamSynchronize(c->stream));
                cudaFree(c->d_bin_entries); c->d_bin_entries = nullptr; c->bin_entries_cap = 0;
                SRT_CUDA(c, cudaMalloc(&c->d_bin_entries, bins * per_bin * 4ull));
                c->bin_entries_cap = bins * per_bin;
            }
            if (bins * c->n_meshes > c->bin_counts_cap) {
                SRT_CUDA(c, cudaStreamSynchronize(c->stream));
                cudaFree(c->d_bin_counts); c->d_bin_counts = nullptr; c->bin_counts_cap = 0;
                SRT_CUDA(c, cudaMalloc(&c->d_bin_counts, bins * c->n_meshes * 4ull));
                c->bin_counts_cap = bins * c->n_meshes;
            }
            a.bin_entries = c->d_bin_entries; a.bin_counts = c->d_bin_counts;
        }
    }

    a.scan_cols = 0; a.scan_chunks = 0; a.hit_head = nullptr; a.hit_next = nullptr;
    const bool big_boxes = box_px * box_py > 64;            // filtered path: the scan's survivors are refined into 16 x 4 sub-boxes
    uint32_t extra_launches = 0;
    if (kernel == SRT_KERNEL_FILTERED) {
        // The scan runs persistent CTAs -- as many as are resident at once -- over the units (chunk, 16-row block, box row,
        // range of 32 box columns); every CTA takes an equal, contiguous share of the units (kernels.cu), so the launch has
        // no partial last wave whatever the frame size, the share of a multi-GPU frame or the residency cap.
        const int sms = c->sms;
        const uint64_t rkey = ((uint64_t)box_px << 48) | ((uint64_t)box_py << 40) | ((uint64_t)scan_k << 32) | c->scan_max_ctas;
        if (c->resident_key != rkey) {                                  // an occupancy query per frame costs microseconds
            c->resident_val = srt::filtered_resident_ctas(box_px, box_py, scan_k, c->scan_max_ctas);
            c->resident_key = rkey;
        }
        int resident = c->resident_val;
        if (resident <= 0) resident = SRT_RESIDENT_CTAS;
        uint64_t chunks = 0;
        for (size_t m = 0; m < c->n_meshes; ++m)
            if ((all || m == 0) && c->meshes[m].n) chunks += c->meshes[m].n_chunks;
        const uint32_t n_cols = (cam.width + box_px - 1) / box_px;
        // unit width: 32 box columns, narrower when the launch is small (big boxes, a share of a multi-GPU frame, a small
        // frame) so that every CTA still gets a dozen units to balance
        uint32_t cols = std::min<uint32_t>(32u, n_cols);
        {
            // cost of the launch = units of the busiest CTA x (boxes of a unit + the unit's fixed overhead, ~3 boxes' worth)
            const uint64_t slots = (uint64_t)sms * (uint64_t)resident;
            double best = 1e300;
            for (uint32_t cw = 64; cw >= 4; cw /= 2) {
                if (cw > n_cols && cw != 4) continue;
                const uint32_t w = std::min(cw, n_cols);
                const uint64_t u = chunks * a.n_blocks * (16 / box_py) * ((n_cols + w - 1) / w);
                const double cost = (double)((u + slots - 1) / slots) * ((double)w + 3.0);
                if (cost < best * 0.995) { best = cost; cols = w; }
            }
        }
        if (c->force_scan_cols > 0) cols = std::min<uint32_t>((uint32_t)c->force_scan_cols, n_cols);
        a.scan_cols = cols; a.scan_parts_x = (n_cols + cols - 1) / cols; a.scan_chunks = (uint32_t)chunks;
        const uint64_t units = chunks * a.n_blocks * (16 / box_py) * a.scan_parts_x;
        a.scan_ctas = (uint32_t)std::min<uint64_t>((uint64_t)sms * (uint64_t)resident, std::max<uint64_t>(units, 1));
        // per-pixel hit lists: a 4-byte head per pixel of the launch's output index range (0xffffffff = empty; the
        // resolve pass leaves them empty again), nodes in the pix list below
        const size_t pixels = (size_t)(parts == 0 ? cam.height : a.n_blocks * 16) * cam.width;
        if (pixels > c->hit_pixels_cap) {
            SRT_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->d_hit_head); c->d_hit_head = nullptr; c->hit_pixels_cap = 0;
            SRT_CUDA(c, cudaMalloc(&c->d_hit_head, pixels * sizeof(uint32_t)));
            SRT_CUDA(c, cudaMemsetAsync(c->d_hit_head, 0xff, pixels * sizeof(uint32_t), c->stream));
            c->hit_pixels_cap = pixels;
        }
        a.hit_head = c->d_hit_head;
        // candidate lists: 8 bytes per survivor of the scan {box, triangle} (3e-4 of the box tests on the dragon) and 12 per
        // survivor of the per-pixel strip tests {pixel, triangle} (~ the hits).  First sized from the scene; the first
        // render of a launch geometry then CHECKS the counts (one synchronisation, below) and grows the lists to twice
        // what that frame needed -- the counts of a scene / camera / launch geometry never change afterwards.
        rc = ensure_lists(c, std::max<uint64_t>(4ull << 20, 32ull * c->total_tris), std::max<uint64_t>(16ull << 20, 64ull * c->total_tris),
                          big_boxes ? std::max<uint64_t>(2ull << 20, 16ull * c->total_tris) : 0);
        if (rc) return rc;
        a.cand = c->d_cand; a.pix = c->d_pix; a.hit_next = c->d_hit_next;
        a.cand_count = c->d_cand_count + 4 * c->cnt_parity; a.cand_count_next = c->d_cand_count + 4 * (c->cnt_parity ^ 1);
        if (big_boxes) { a.big = c->d_big; a.big_count = a.cand_count + 2; a.big_cap = c->big_cap; }
        a.cand_cap = c->cand_cap; a.pix_cap = c->pix_cap; a.scan_max_ctas = c->scan_max_ctas; a.sm_count = (uint32_t)sms;
        // the blocks of a CTA's contiguous range of units are spread over the whole frame (kernels.cu): stride ~ 0.618 n, coprime to n
        {
            uint32_t st = std::max<uint32_t>(1u, (uint32_t)((double)a.n_blocks * 0.6180339887));
            auto gcd = [](uint32_t x, uint32_t y) { while (y) { const uint32_t t = x % y; x = y; y = t; } return x; };
            while (st > 1 && gcd(st, a.n_blocks) != 1) --st;
            a.scan_blk_stride = a.n_blocks ? st : 1;
        }
        if (c->force_cand_cap) { a.cand_cap = std::min(a.cand_cap, c->force_cand_cap); a.big_cap = std::min(a.big_cap, c->force_cand_cap); }
        if (c->force_pix_cap) a.pix_cap = std::min(a.pix_cap, c->force_pix_cap);
        if (c->n_meshes > 32 || c->total_tris >= (1ull << 27)) kernel = SRT_KERNEL_EXACT;     // candidate entries: 5-bit mesh, 27-bit triangle
    }
    const bool prof = c->profiling;
    c->verified_last_render = repack;
    if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if (kernel == SRT_KERNEL_FILTERED) {
        // launch geometry the list sizes were verified for (scene / camera changes reset it through `packed`)
        const uint64_t key = ((uint64_t)row_begin << 40) ^ ((uint64_t)row_end << 20) ^ ((uint64_t)parts << 12) ^ ((uint64_t)part << 4) ^
                             ((uint64_t)mode << 2) ^ (all ? 2u : 0u) ^ ((uint64_t)box_px << 52) ^ ((uint64_t)box_py << 58) ^ 1u;
        const bool forced = c->force_cand_cap || c->force_pix_cap;              // tests exercise the overflow fallback
        const bool verify = !forced && (repack || c->lists_key != key);
        c->verified_last_render = verify;
        // does launch_lbuffer_filtered launch anything?  (an empty launch touches no counter: the pairs keep their roles)
        const bool launches = !(a.row_end <= a.row_begin || a.cam.W == 0 || a.n_blocks == 0 || a.scan_chunks == 0);
        if (!verify && !forced) { a.cand_hint = c->verified_cand; a.pix_hint = c->verified_pix; a.big_hint = c->verified_big; }
        for (int attempt = 0;; ++attempt) {
            // the pair of counters this launch counts into was left at zero by the previous scan (or at allocation);
            // a repeated attempt counts into the same pair again and clears it first
            if (attempt > 0) SRT_CUDA(c, cudaMemsetAsync(a.cand_count, 0, 4 * sizeof(uint32_t), c->stream));
            if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
            SRT_CUDA(c, srt::launch_lbuffer_filtered(a, c->stream));
            if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
            SRT_CUDA(c, srt::launch_filtered_candidates(a, c->stream));
            if (!verify) break;
            // First render of this launch geometry: wait for the counts.  A list that overflowed would make the resolve pass
            // recompute every pixel with the exact sequence (correct, but hours on a large scene): grow and redo instead.
            SRT_CUDA(c, cudaMemcpyAsync(c->h_cand_count, a.cand_count, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            SRT_CUDA(c, cudaStreamSynchronize(c->stream));
            const uint64_t n_cand = launches ? c->h_cand_count[0] : 0, n_pix = launches ? c->h_cand_count[1] : 0;
            const uint64_t n_big = (launches && big_boxes) ? c->h_cand_count[2] : 0;
            if (n_cand <= a.cand_cap && n_pix <= a.pix_cap && n_big <= a.big_cap) {
                c->lists_key = key;
                c->verified_cand = (uint32_t)std::max<uint64_t>(n_cand, 1); c->verified_pix = (uint32_t)std::max<uint64_t>(n_pix, 1);
                c->verified_big = (uint32_t)std::max<uint64_t>(n_big, 1);
                break;
            }
            if (attempt >= 4) break;                                   // give up growing: the resolve pass stays exact
            // the counts behind a truncated list are lower bounds: grow generously, the next attempt measures them
            const bool cut = n_big > a.big_cap;
            rc = ensure_lists(c, cut ? std::max<uint64_t>(4 * n_cand, 2 * n_big) : 2 * n_cand,
                              std::max<uint64_t>(2 * n_pix, (n_cand <= a.cand_cap && !cut) ? 0 : 4 * n_pix), big_boxes ? 2 * n_big : 0);
            if (rc) return rc;
            // the exact-test stage has appended hits to the pixels' lists: empty them again before the next attempt
            {
                const size_t px = (size_t)(parts == 0 ? cam.height : a.n_blocks * 16) * cam.width;
                SRT_CUDA(c, cudaMemsetAsync(c->d_hit_head, 0xff, px * sizeof(uint32_t), c->stream));
            }
            a.cand = c->d_cand; a.pix = c->d_pix; a.hit_next = c->d_hit_next; a.cand_cap = c->cand_cap; a.pix_cap = c->pix_cap;
            if (big_boxes) { a.big = c->d_big; a.big_cap = c->big_cap; }
        }
        if (launches) c->cnt_parity ^= 1;                              // the scan has cleared the other pair
        SRT_CUDA(c, srt::launch_resolve_hits(a, c->stream)); extra_launches = big_boxes ? 4 : 3;   // + (refine,) candidates, exact tests
    } else if (kernel == SRT_KERNEL_CULLED) {
        SRT_CUDA(c, srt::launch_lbuffer_bins(a, c->stream)); extra_launches = 1;
        if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
        SRT_CUDA(c, srt::launch_lbuffer_culled(a, c->stream));
        if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    } else {
        if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
        SRT_CUDA(c, srt::launch_lbuffer_exact(a, audit, c->stream));
        if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    }
    if (prof) SRT_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    c->timing.kernel_used = kernel;
    c->timing.scan_cols = a.scan_cols;
    c->timing.launches = ((band_rows > 0) ? 1 + extra_launches : 0) + (repack ? 1 + (uint32_t)c->n_meshes : 0);
    const uint64_t tris = all ? c->total_tris : c->mesh0_tris;
    c->timing.pair_tests = (uint64_t)band_rows * cam.width * tris;
    c->timing.exact_tests = 0; c->timing.filter_violations = 0;
    c->counters_pending = a.counters != nullptr;
    c->timing_pending = prof || c->counters_pending;
    if (!prof) { c->timing.render_ms = 0.0f; c->timing.main_kernel_ms = 0.0f; }
    return SRT_OK;
}

} // namespace

// =============================================================================
extern "C" {

int srt_version(void) { return SRT_VERSION; }

const char* srt_last_error(const srt_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

void srt_default_params(srt_params* p)
{
    if (!p) return;
    p->incident = 80.000f;          // lbuffer.cxx:314
    p->length_to_cm = 0.1;          // lbuffer.cxx:808
    p->repair_depth = 4;            // lbuffer.cxx:339
    p->kernel = SRT_KERNEL_AUTO;
    p->flags = 0;
    p->pixels_per_thread = 0;
}

int srt_create(srt_ctx** out, int device_id)
{
    if (!out) return fail(nullptr, SRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, SRT_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n) return fail(nullptr, SRT_ERR_INVALID, "device %d out of range [0, %d)", device_id, n);
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess)
        return fail(nullptr, SRT_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SRT_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device_id, prop.major, prop.minor);
    srt_ctx* c = new (std::nothrow) srt_ctx();
    if (!c) return fail(nullptr, SRT_ERR_NOMEM, "out of host memory");
    c->device = device_id;
    c->sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if ((e = cudaSetDevice(device_id)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c; return fail(nullptr, SRT_ERR_CUDA, "stream creation: %s", cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    for (auto& ev : c->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) { srt_destroy(c); return fail(nullptr, SRT_ERR_CUDA, "event creation: %s", cudaGetErrorString(e)); }
    *out = c;
    return SRT_OK;
}

void srt_destroy(srt_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_scene(c); free_frame(c); free_stage(c);
    for (auto& g : c->frame_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    c->frame_graphs.clear();
    cudaFree(c->d_barrier_ctr);
    cudaFree(c->d_counters);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int srt_set_stream(srt_ctx* c, void* s)
{
    if (!c) return SRT_ERR_INVALID;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return SRT_OK;
}

int srt_synchronize(srt_ctx* c)
{
    if (!c) return SRT_ERR_INVALID;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return SRT_OK;
}

int srt_set_profiling(srt_ctx* c, int enable)
{
    if (!c) return SRT_ERR_INVALID;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->profiling = enable != 0;
    c->timing_pending = false;
    return SRT_OK;
}

int srt_set_scan_residency(srt_ctx* c, uint32_t max_ctas_per_sm)
{
    if (!c) return SRT_ERR_INVALID;
    if (max_ctas_per_sm != 0 && max_ctas_per_sm < 4) return fail(c, SRT_ERR_INVALID, "the scan can be capped at 4 or more CTAs per SM (0 = no cap)");
    c->scan_max_ctas = max_ctas_per_sm;
    return SRT_OK;
}

int srt_set_host_copy_threads(srt_ctx* c, int threads)
{
    if (!c) return SRT_ERR_INVALID;
    if (threads < -1 || threads > 32) return fail(c, SRT_ERR_INVALID, "copy threads: -1 (default), 0 (off) ... 32");
    c->copy_threads = threads;
    return SRT_OK;
}

int srt_debug_tune_host_copy(srt_ctx* c, size_t chunk_bytes, int spin_us)
{
    if (!c) return SRT_ERR_INVALID;
    if (chunk_bytes != 0 && (chunk_bytes < (64u << 10) || chunk_bytes > kStageBytes || (chunk_bytes & (chunk_bytes - 1)) != 0))
        return fail(c, SRT_ERR_INVALID, "copy chunk: 0 (default) or a power of two from 64 KiB to 32 MiB");
    if (spin_us < -1 || spin_us > 1000000) return fail(c, SRT_ERR_INVALID, "spin time: -1 (default), 0 ... 1000000 us");
    c->copy_chunk = chunk_bytes;
    c->copy_spin_us = spin_us;
    return SRT_OK;
}

int srt_debug_host_copy(void* dst, const void* src, size_t bytes, size_t chunk_bytes, int threads, int spin_us)
{
    if ((!dst || !src) && bytes) return SRT_ERR_INVALID;
    if (chunk_bytes == 0 || threads < 0 || threads > 32 || spin_us < 0 || spin_us > 1000000) return SRT_ERR_INVALID;
    srt::HostCopyPool pool(threads, spin_us > 0 ? (unsigned)spin_us : 0u);
    pool.activate();
    for (size_t off = 0; off < bytes; off += chunk_bytes)
        pool.submit((char*)dst + off, (const char*)src + off, std::min(chunk_bytes, bytes - off));
    pool.drain();
    return SRT_OK;
}

int srt_debug_host_copy_stats(srt_ctx* c, uint64_t* staged_copies)
{
    if (!c || !staged_copies) return SRT_ERR_INVALID;
    *staged_copies = c->staged_copies;
    return SRT_OK;
}

// ---- host helpers -----------------------------------------------------------
int srt_triangle_normals(const float* vertices, uint32_t n, float* out)
{
    if ((!vertices || !out) && n) return SRT_ERR_INVALID;
    for (uint32_t i = 0; i < n; ++i) srt::compute_normal(vertices + 9 * (size_t)i, out + 3 * (size_t)i);
    return SRT_OK;
}

int srt_scene_bbox(const srt_mesh_desc* meshes, int n_meshes, float lo[3], float hi[3])
{
    if (!lo || !hi || n_meshes < 0 || (n_meshes && !meshes)) return SRT_ERR_INVALID;
    const float inf = std::numeric_limits<float>::infinity();
    for (int k = 0; k < 3; ++k) { lo[k] = inf; hi[k] = -inf; }
    for (int m = 0; m < n_meshes; ++m) {
        if (meshes[m].n_triangles && !meshes[m].vertices) return SRT_ERR_INVALID;
        // TriangleMesh::computeBoundingBox (TriangleMesh.cxx:192-228): std::min / std::max over P1, P2, P3
        float mlo[3] = { inf, inf, inf }, mhi[3] = { -inf, -inf, -inf };
        for (size_t i = 0; i < (size_t)meshes[m].n_triangles * 3; ++i)
            for (int k = 0; k < 3; ++k) {
                const float v = meshes[m].vertices[i * 3 + k];
                mlo[k] = (v < mlo[k]) ? v : mlo[k];
                mhi[k] = (mhi[k] < v) ? v : mhi[k];
            }
        for (int k = 0; k < 3; ++k) {       // getBBox (lbuffer.cxx:585-608)
            lo[k] = (mlo[k] < lo[k]) ? mlo[k] : lo[k];
            hi[k] = (hi[k] < mhi[k]) ? mhi[k] : hi[k];
        }
    }
    return SRT_OK;
}

int srt_camera_from_bbox(const float lo[3], const float hi[3], uint32_t W, uint32_t H, srt_camera* out)
{
    if (!lo || !hi || !out || W == 0 || H == 0) return SRT_ERR_INVALID;
    // initialiseRayTracing, lbuffer.cxx:611-667.  Float ops, except where the reference's
    // Vec3::operator/(double) and `diagonal * 0.6` promote to double.
    float range[3], centre[3];
    for (int k = 0; k < 3; ++k) {
        range[k] = hi[k] - lo[k];                                   // :620
        centre[k] = lo[k] + (float)((double)range[k] / 2.0);        // :621
    }
    const float diagonal = std::sqrt(range[0] * range[0] + range[1] * range[1] + range[2] * range[2]);   // :631
    const float up[3] = { 0.0f, 0.0f, -1.0f };                      // :633
    float origin[3] = { centre[0] - diagonal * 1, centre[1] - 0.0f, centre[2] - 0.0f };                     // :635
    float detector[3] = { centre[0] + (float)((double)diagonal * 0.6), centre[1] + 0.0f, centre[2] + 0.0f }; // :636
    float d[3] = { detector[0] - origin[0], detector[1] - origin[1], detector[2] - origin[2] };             // :638
    for (int pass = 0; pass < 2; ++pass) {                          // normalize() :639 and normalise() :648
        const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        d[0] /= len; d[1] /= len; d[2] /= len;
    }
    // right = direction x up, :649 (Vec3::crossProduct)
    const float right[3] = { d[1] * up[2] - d[2] * up[1], d[2] * up[0] - d[0] * up[2], d[0] * up[1] - d[1] * up[0] };
    // pixel spacing, :739-741
    const float res1 = range[2] / (float)W, res2 = range[1] / (float)H;
    const float spacing = 2 * ((res1 < res2) ? res2 : res1);
    for (int k = 0; k < 3; ++k) { out->origin[k] = origin[k]; out->detector[k] = detector[k]; out->up[k] = up[k]; out->right[k] = right[k]; }
    out->pixel_spacing[0] = spacing; out->pixel_spacing[1] = spacing;
    out->width = W; out->height = H;
    return SRT_OK;
}

int srt_partition_rows(uint32_t H, uint32_t n_parts, uint32_t part, uint32_t* rb, uint32_t* re)
{
    if (!rb || !re || n_parts == 0 || part >= n_parts) return SRT_ERR_INVALID;
    // total / n each, the first total % n parts get one more (lbuffer.cxx:701-719, rows variant :312-334)
    const uint32_t per = H / n_parts, rem = H % n_parts;
    *rb = part * per + (part < rem ? part : rem);
    *re = *rb + per + (part < rem ? 1u : 0u);
    return SRT_OK;
}

int srt_host_alloc(void** out, size_t bytes)
{
    if (!out) return SRT_ERR_INVALID;
    *out = nullptr;
    if (bytes == 0) return SRT_OK;
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { *out = nullptr; return fail(nullptr, e == cudaErrorMemoryAllocation ? SRT_ERR_NOMEM : SRT_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e)); }
    return SRT_OK;
}

void srt_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- JPEG output (nvJPEG: the encoder runs on the GPU; replaces the libjpeg writers of src/Image.cxx:85-146 and
// src/Image.bakk.cxx:170-226 -- quality 100, no chroma subsampling) ---------------------------------
int srt_write_jpeg(srt_ctx* c, const char* path, const uint8_t* rgb, uint32_t width, uint32_t height, int quality)
{
    if (!c) return SRT_ERR_INVALID;
    if (!path || !rgb || width == 0 || height == 0 || quality < 1 || quality > 100)
        return fail(c, SRT_ERR_INVALID, "srt_write_jpeg: bad arguments");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)width * height * 3;
    if (bytes > c->rgb_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_rgb); c->d_rgb = nullptr; c->rgb_cap = 0;
        SRT_CUDA(c, cudaMalloc(&c->d_rgb, bytes)); c->rgb_cap = bytes;
    }
    SRT_CUDA(c, cudaMemcpyAsync(c->d_rgb, rgb, bytes, cudaMemcpyHostToDevice, c->stream));
    nvjpegHandle_t h = nullptr; nvjpegEncoderState_t st = nullptr; nvjpegEncoderParams_t ep = nullptr;
    auto cleanup = [&]() {
        if (ep) nvjpegEncoderParamsDestroy(ep);
        if (st) nvjpegEncoderStateDestroy(st);
        if (h) nvjpegDestroy(h);
    };
#define SRT_NVJPEG(call) do { const nvjpegStatus_t s__ = (call); if (s__ != NVJPEG_STATUS_SUCCESS) { cleanup(); \
        return fail(c, SRT_ERR_CUDA, "%s: nvjpeg status %d", #call, (int)s__); } } while (0)
    SRT_NVJPEG(nvjpegCreateSimple(&h));
    SRT_NVJPEG(nvjpegEncoderStateCreate(h, &st, c->stream));
    SRT_NVJPEG(nvjpegEncoderParamsCreate(h, &ep, c->stream));
    SRT_NVJPEG(nvjpegEncoderParamsSetQuality(ep, quality, c->stream));
    SRT_NVJPEG(nvjpegEncoderParamsSetSamplingFactors(ep, NVJPEG_CSS_444, c->stream));
    nvjpegImage_t img{};
    img.channel[0] = c->d_rgb; img.pitch[0] = (size_t)width * 3;
    SRT_NVJPEG(nvjpegEncodeImage(h, st, ep, &img, NVJPEG_INPUT_RGBI, (int)width, (int)height, c->stream));
    size_t len = 0;
    SRT_NVJPEG(nvjpegEncodeRetrieveBitstream(h, st, nullptr, &len, c->stream));
    std::vector<unsigned char> buf(len);
    SRT_NVJPEG(nvjpegEncodeRetrieveBitstream(h, st, buf.data(), &len, c->stream));
#undef SRT_NVJPEG
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) return fail(c, SRT_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return fail(c, SRT_ERR_INVALID, "Cannot create the file %s", path);
    const size_t w = std::fwrite(buf.data(), 1, len, f);
    std::fclose(f);
    if (w != len) return fail(c, SRT_ERR_INVALID, "short write to %s", path);
    return SRT_OK;
}

// ---- scene --------------------------------------------------------------------
int srt_set_meshes(srt_ctx* c, int n_meshes, const srt_mesh_desc* meshes)
{
    if (!c) return SRT_ERR_INVALID;
    if (n_meshes <= 0 || !meshes) return fail(c, SRT_ERR_INVALID, "need at least one mesh");
    for (int m = 0; m < n_meshes; ++m)
        if (meshes[m].n_triangles && !meshes[m].vertices)
            return fail(c, SRT_ERR_INVALID, "mesh %d: vertices is NULL (buffer size error)", m);
    SRT_CUDA(c, cudaSetDevice(c->device));
    if (c->meshes.size() < (size_t)n_meshes) c->meshes.resize((size_t)n_meshes);
    c->packed = false;
    ++c->scene_serial;
    c->n_meshes = 0;
    for (int m = 0; m < n_meshes; ++m) {
        const srt_mesh_desc& d = meshes[m];
        MeshBuf& mb = c->meshes[m];
        int rc = ensure_mesh(c, mb, d.n_triangles, d.normals != nullptr);
        if (rc) return rc;
        mb.n = d.n_triangles;
        mb.n_padded = (d.n_triangles + srt::kPadTris - 1) / srt::kPadTris * srt::kPadTris;
        mb.mu = d.mu;
        mb.has_normals = d.normals != nullptr;
        mb.material = d.material;
        mb.has_texcoords = d.texcoords != nullptr && d.n_triangles > 0;
        if (mb.has_texcoords) {
            if (d.n_triangles > mb.texcoord_cap) {
                SRT_CUDA(c, cudaStreamSynchronize(c->stream));
                cudaFree(mb.d_texcoords); mb.d_texcoords = nullptr;
                SRT_CUDA(c, cudaMalloc(&mb.d_texcoords, (size_t)d.n_triangles * 9 * sizeof(float))); mb.texcoord_cap = d.n_triangles;
            }
            SRT_CUDA(c, cudaMemcpyAsync(mb.d_texcoords, d.texcoords, (size_t)d.n_triangles * 9 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        mb.tex_w = mb.tex_h = 0;
        if (d.texture_rgb && d.tex_w && d.tex_h) {
            const size_t bytes = (size_t)d.tex_w * d.tex_h * 3;
            if (bytes > mb.tex_cap) {
                SRT_CUDA(c, cudaStreamSynchronize(c->stream));
                cudaFree(mb.d_tex); mb.d_tex = nullptr;
                SRT_CUDA(c, cudaMalloc(&mb.d_tex, bytes)); mb.tex_cap = bytes;
            }
            SRT_CUDA(c, cudaMemcpyAsync(mb.d_tex, d.texture_rgb, bytes, cudaMemcpyHostToDevice, c->stream));
            mb.tex_w = d.tex_w; mb.tex_h = d.tex_h;
        }
        // pageable source: the runtime stages the data before returning, the caller's buffer is free again
        if (d.n_triangles) {
            SRT_CUDA(c, cudaMemcpyAsync(mb.d_vertices, d.vertices, (size_t)d.n_triangles * 9 * sizeof(float),
                                        cudaMemcpyHostToDevice, c->stream));
            if (d.normals)
                SRT_CUDA(c, cudaMemcpyAsync(mb.d_normals, d.normals, (size_t)d.n_triangles * 3 * sizeof(float),
                                            cudaMemcpyHostToDevice, c->stream));
        }
    }
    c->n_meshes = (size_t)n_meshes;
    return SRT_OK;
}

int srt_set_camera(srt_ctx* c, const srt_camera* cam)
{
    if (!c) return SRT_ERR_INVALID;
    if (!cam) return fail(c, SRT_ERR_INVALID, "camera is NULL");
    if (cam->width == 0 || cam->height == 0 || (uint64_t)cam->width * cam->height >= (1ull << 32))
        return fail(c, SRT_ERR_INVALID, "image size %u x %u not supported", cam->width, cam->height);
    c->camera = *cam;
    c->have_camera = true;
    c->packed = false;
    return SRT_OK;
}

// ---- L-buffer -------------------------------------------------------------------
int srt_render_lbuffer_device(srt_ctx* c, const srt_params* p, uint32_t rb, uint32_t re,
                              float* dL, float* dd, int32_t* ds, uint32_t* dh)
{
    if (!c) return SRT_ERR_INVALID;
    SRT_CUDA(c, cudaSetDevice(c->device));
    return render_device(c, p, rb, re, dL, dd, ds, dh);
}

int srt_interleave_slot_rows(uint32_t height, uint32_t n_parts, uint32_t* slot_rows)
{
    if (!slot_rows || n_parts == 0) return SRT_ERR_INVALID;
    const uint32_t blocks = (height + 15) / 16;
    *slot_rows = ((blocks + n_parts - 1) / n_parts) * 16;
    return SRT_OK;
}

int srt_render_lbuffer_interleaved_device(srt_ctx* c, const srt_params* p, uint32_t n_parts, uint32_t part,
                                          float* d_slot, float* dd, int32_t* ds, uint32_t* dh)
{
    if (!c) return SRT_ERR_INVALID;
    if (n_parts == 0 || part >= n_parts) return fail(c, SRT_ERR_INVALID, "part %u of %u", part, n_parts);
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    return render_device(c, p, 0, c->camera.height, d_slot, dd, ds, dh, n_parts, part);
}

int srt_render_lbuffer_blocks_device(srt_ctx* c, const srt_params* p, uint32_t n_parts, uint32_t part,
                                     uint32_t n_frames, float* const* d_frames)
{
    if (!c) return SRT_ERR_INVALID;
    if (n_parts == 0 || part >= n_parts) return fail(c, SRT_ERR_INVALID, "part %u of %u", part, n_parts);
    if (n_frames == 0 || n_frames > 8 || !d_frames) return fail(c, SRT_ERR_INVALID, "1 to 8 frame pointers expected");
    for (uint32_t i = 0; i < n_frames; ++i)
        if (!d_frames[i]) return fail(c, SRT_ERR_INVALID, "frame pointer %u is NULL", i);
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    // a single pointer with SRT_FLAG_FRAME0_MULTICAST is a multicast mapping; n_parts pointers (own + every peer) with
    // SRT_FLAG_L_HALO_ONLY send only the repair stencil's halo rows, and only to the two neighbouring ranks
    const bool multicast = (p && (p->flags & SRT_FLAG_FRAME0_MULTICAST) != 0) && n_frames == 1;
    const bool halo = (p && (p->flags & SRT_FLAG_L_HALO_ONLY) != 0) && n_frames == n_parts && n_parts > 1;
    return render_device(c, p, 0, c->camera.height, d_frames[0], nullptr, nullptr, nullptr, n_parts, part, 0,
                         d_frames + 1, n_frames - 1, /*image_positions*/ true, multicast, halo);
}

int srt_deinterleave_device(srt_ctx* c, uint32_t n_parts, const float* d_slots, float* d_frame)
{
    if (!c) return SRT_ERR_INVALID;
    if (n_parts == 0 || !d_slots || !d_frame || d_slots == d_frame) return fail(c, SRT_ERR_INVALID, "bad deinterleave arguments");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    uint32_t slot_rows = 0;
    srt_interleave_slot_rows(c->camera.height, n_parts, &slot_rows);
    SRT_CUDA(c, srt::launch_deinterleave(d_slots, d_frame, c->camera.width, c->camera.height, n_parts, slot_rows, c->stream));
    return SRT_OK;
}

int srt_peer_barrier_device(srt_ctx* c, uint32_t n_ranks, uint32_t rank, uint32_t* const* d_flags, uint32_t epoch, uint32_t phases)
{
    if (!c) return SRT_ERR_INVALID;
    if (phases == 0 || phases > 3) return fail(c, SRT_ERR_INVALID, "phases: 1 = signal, 2 = wait, 3 = both");
    if (n_ranks == 0 || n_ranks > 8 || rank >= n_ranks || !d_flags) return fail(c, SRT_ERR_INVALID, "rank %u of %u (1 to 8 ranks, flag arrays required)", rank, n_ranks);
    srt::PeerFlags f{};
    for (uint32_t j = 0; j < n_ranks; ++j) {
        if (!d_flags[j]) return fail(c, SRT_ERR_INVALID, "flag array %u is NULL", j);
        f.p[j] = d_flags[j];
    }
    SRT_CUDA(c, cudaSetDevice(c->device));
    SRT_CUDA(c, srt::launch_peer_barrier(f, n_ranks, rank, epoch, phases, nullptr, c->stream));
    return SRT_OK;
}

int srt_repair_blocks_device(srt_ctx* c, const srt_params* p, uint32_t n_parts, uint32_t part, const float* d_L,
                             uint32_t n_images, float* const* d_images)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !d_L) return fail(c, SRT_ERR_INVALID, "params / L frame is NULL");
    if (n_parts == 0 || part >= n_parts) return fail(c, SRT_ERR_INVALID, "part %u of %u", part, n_parts);
    if (n_images == 0 || n_images > 8 || !d_images) return fail(c, SRT_ERR_INVALID, "1 to 8 image pointers expected");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    const bool multicast = (p->flags & SRT_FLAG_FRAME0_MULTICAST) != 0 && n_images == 1;
    srt::RepairOut out{};
    bool aligned = (c->camera.width & 3u) == 0;
    for (uint32_t i = 0; i < n_images; ++i) {
        if (!d_images[i] || d_images[i] == d_L) return fail(c, SRT_ERR_INVALID, "image pointer %u is NULL or aliases the L frame", i);
        aligned = aligned && ((uintptr_t)d_images[i] & 15u) == 0;
        if (i == 0) out.image = d_images[0]; else out.extra[i - 1] = d_images[i];
    }
    out.n_extra = n_images - 1; out.multicast = multicast ? 1 : 0; out.vec4 = aligned ? 1 : 0;
    SRT_CUDA(c, cudaSetDevice(c->device));
    if (c->profiling) SRT_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    SRT_CUDA(c, srt::launch_repair_blocks(d_L, out, c->camera.width, c->camera.height, p->repair_depth, n_parts, part, c->stream));
    if (c->profiling) SRT_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    return SRT_OK;
}

} // extern "C"

namespace {

// The launches of one sharded frame: render of the rank's blocks -> barrier -> sharded repair -> barrier, the barriers taking
// their epochs from the context's device counter.
int frame_sequence(srt_ctx* c, const srt_params* p, uint32_t n_parts, uint32_t part, uint32_t n_frames, float* const* d_frames,
                   uint32_t* const* d_flags, const float* d_own_frame, uint32_t n_images, float* const* d_images)
{
    int rc = srt_render_lbuffer_blocks_device(c, p, n_parts, part, n_frames, d_frames);
    if (rc) return rc;
    srt::PeerFlags f{};
    for (uint32_t j = 0; j < n_parts && n_parts > 1; ++j) f.p[j] = d_flags[j];
    // barrier 1: every rank's rows have landed in this rank's L frame
    if (n_parts > 1) SRT_CUDA(c, srt::launch_peer_barrier(f, n_parts, part, 0, 3, c->d_barrier_ctr, c->stream));
    rc = srt_repair_blocks_device(c, p, n_parts, part, d_own_frame, n_images, d_images);
    if (rc) return rc;
    // barrier 2: every rank's repaired rows have landed in this rank's image
    if (n_parts > 1) SRT_CUDA(c, srt::launch_peer_barrier(f, n_parts, part, 0, 3, c->d_barrier_ctr, c->stream));
    return SRT_OK;
}

} // namespace

extern "C" {

int srt_render_xray_blocks_device(srt_ctx* c, const srt_params* p, uint32_t n_parts, uint32_t part, uint32_t n_frames,
                                  float* const* d_frames, uint32_t* const* d_flags, uint32_t epoch, const float* d_own_frame,
                                  uint32_t n_images, float* const* d_images)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !d_own_frame) return fail(c, SRT_ERR_INVALID, "params / own frame is NULL");
    if (epoch == 0 || epoch >= 0x7fffffffu) return fail(c, SRT_ERR_INVALID, "epoch counts the frames of this flag set from 1");
    if (n_parts == 0 || n_parts > 8 || part >= n_parts) return fail(c, SRT_ERR_INVALID, "part %u of %u (1 to 8 parts)", part, n_parts);
    if (n_frames == 0 || n_frames > 8 || !d_frames || n_images == 0 || n_images > 8 || !d_images)
        return fail(c, SRT_ERR_INVALID, "1 to 8 frame and image pointers expected");
    if (n_parts > 1) {
        if (!d_flags) return fail(c, SRT_ERR_INVALID, "flag arrays required");
        for (uint32_t j = 0; j < n_parts; ++j) if (!d_flags[j]) return fail(c, SRT_ERR_INVALID, "flag array %u is NULL", j);
    }
    SRT_CUDA(c, cudaSetDevice(c->device));
    // the barriers count their epochs (2 e - 1, 2 e) in device memory; the counter follows the caller's numbering
    if (!c->d_barrier_ctr) {
        SRT_CUDA(c, cudaMalloc(&c->d_barrier_ctr, sizeof(uint32_t)));
        SRT_CUDA(c, cudaMemsetAsync(c->d_barrier_ctr, 0, sizeof(uint32_t), c->stream));
        c->barrier_epochs = 0;
    }
    const uint32_t before = 2 * epoch - 2;
    if (c->barrier_epochs != before) {                     // a new numbering (another set of flag arrays, a restart)
        SRT_CUDA(c, cudaMemcpyAsync(c->d_barrier_ctr, &before, sizeof before, cudaMemcpyHostToDevice, c->stream));
        c->barrier_epochs = before;
    }

    // ---- the call's signature: everything its launches depend on
    std::vector<uint64_t> sig;
    sig.reserve(40);
    auto put = [&](uint64_t v) { sig.push_back(v); };
    uint64_t bits = 0;
    std::memcpy(&bits, &p->incident, sizeof(float)); put(bits);
    std::memcpy(&bits, &p->length_to_cm, sizeof(double)); put(bits);
    put(((uint64_t)p->repair_depth << 32) | p->kernel); put(((uint64_t)p->flags << 32) | p->pixels_per_thread);
    put(((uint64_t)n_parts << 48) | ((uint64_t)part << 32) | ((uint64_t)n_frames << 16) | n_images);
    for (uint32_t i = 0; i < n_frames; ++i) put((uint64_t)(uintptr_t)d_frames[i]);
    for (uint32_t i = 0; i < n_images; ++i) put((uint64_t)(uintptr_t)d_images[i]);
    for (uint32_t j = 0; j < n_parts && n_parts > 1; ++j) put((uint64_t)(uintptr_t)d_flags[j]);
    put((uint64_t)(uintptr_t)d_own_frame); put((uint64_t)(uintptr_t)c->stream);
    put(c->pack_serial); put(((uint64_t)c->cnt_parity << 32) | c->scan_max_ctas);
    static const bool graphs_off = std::getenv("SRT_NO_GRAPH") != nullptr;
    const bool forced = c->force_cand_cap || c->force_pix_cap || c->force_bin_cap || c->force_scan_cols;
    const bool may_graph = !graphs_off && c->packed && !c->profiling && !c->count_exact && !forced;

    srt_ctx::FrameGraph* g = nullptr;
    for (auto& e : c->frame_graphs) if (e.sig == sig) { g = &e; break; }
    if (may_graph && g && g->exec && g->lists_key_after == c->lists_key) {
        // replay: the graph holds a scan that counts into this parity's counters and clears the other set
        SRT_CUDA(c, cudaGraphLaunch(g->exec, c->stream));
        if (g->flips_parity) c->cnt_parity ^= 1;
        c->barrier_epochs += 2;
        c->timing_pending = false;
        return SRT_OK;
    }
    if (may_graph && g && !g->exec && g->steady_runs >= 1 && g->lists_key_after == c->lists_key) {
        // the signature has run in the steady state before (no packing, no list verification, no allocation): capture it
        const int parity0 = c->cnt_parity;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc_seq = frame_sequence(c, p, n_parts, part, n_frames, d_frames, d_flags, d_own_frame, n_images, d_images);
            const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
            const bool clean = rc_seq == SRT_OK && ce == cudaSuccess && graph && !c->verified_last_render;
            c->cnt_parity = parity0;                       // nothing has run yet
            if (clean && cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) g->exec = nullptr;
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (clean && g->exec) {
                SRT_CUDA(c, cudaGraphLaunch(g->exec, c->stream));
                if (g->flips_parity) c->cnt_parity ^= 1;
                c->barrier_epochs += 2;
                c->timing_pending = false;
                return SRT_OK;
            }
            g->steady_runs = 0;                            // not capturable as it is: run it normally below
        }
    }
    // ---- plain launches
    const uint64_t lists_before = c->lists_key;
    const int parity_before = c->cnt_parity;
    const int rc = frame_sequence(c, p, n_parts, part, n_frames, d_frames, d_flags, d_own_frame, n_images, d_images);
    if (rc) return rc;
    c->barrier_epochs += 2;
    if (may_graph) {
        if (!g) {
            if (c->frame_graphs.size() >= 16) {            // bounded cache: drop the oldest entry
                if (c->frame_graphs.front().exec) cudaGraphExecDestroy(c->frame_graphs.front().exec);
                c->frame_graphs.erase(c->frame_graphs.begin());
            }
            c->frame_graphs.emplace_back();
            g = &c->frame_graphs.back();
            g->sig = sig;
        }
        const bool steady = !c->verified_last_render && lists_before == c->lists_key;
        g->lists_key_after = c->lists_key;                 // a replay is valid while no other launch geometry has been verified since
        g->flips_parity = parity_before != c->cnt_parity;  // the filtered path ran: its scan alternates between the two counter sets
        g->steady_runs = steady ? g->steady_runs + 1 : 0;
    }
    return SRT_OK;
}

int srt_render_lbuffer(srt_ctx* c, const srt_params* p, uint32_t rb, uint32_t re,
                       float* L_out, float* dist_out, int32_t* ssum_out, uint32_t* hits_out)
{
    if (!c) return SRT_ERR_INVALID;
    if (!L_out) return fail(c, SRT_ERR_INVALID, "L_out is NULL");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const size_t W = c->camera.width, pixels = W * c->camera.height;
    int rc = ensure_frame(c, pixels);
    if (rc) return rc;
    rc = render_device(c, p, rb, re, c->d_L, dist_out ? c->d_dist : nullptr, ssum_out ? c->d_ssum : nullptr,
                       hits_out ? c->d_hits : nullptr);
    if (rc) return rc;
    const size_t off = (size_t)rb * W, cnt = (size_t)(re - rb) * W;
    if (cnt) {
        if ((rc = copy_to_host(c, L_out + off, c->d_L + off, cnt * 4)) != 0) return rc;
        if (dist_out && (rc = copy_to_host(c, dist_out + off, c->d_dist + off, cnt * 4)) != 0) return rc;
        if (ssum_out && (rc = copy_to_host(c, ssum_out + off, c->d_ssum + off, cnt * 4)) != 0) return rc;
        if (hits_out && (rc = copy_to_host(c, hits_out + off, c->d_hits + off, cnt * 4)) != 0) return rc;
    }
    if (c->profiling) SRT_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (int erc = elapsed_ms(c, &c->timing.render_ms, 0, 1)) return erc;
    if (int erc = elapsed_ms(c, &c->timing.d2h_ms, 1, 2)) return erc;
    c->timing.repair_ms = 0.0f; c->timing.h2d_ms = 0.0f;
    return SRT_OK;
}

// ---- repair -----------------------------------------------------------------------
int srt_repair_device(srt_ctx* c, const srt_params* p, const float* dL, float* dimg)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !dL || !dimg || dL == dimg) return fail(c, SRT_ERR_INVALID, "repair needs distinct non-NULL buffers");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    if (c->profiling) SRT_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    SRT_CUDA(c, srt::launch_repair(dL, dimg, c->camera.width, c->camera.height, p->repair_depth, c->stream));
    if (c->profiling) SRT_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    return SRT_OK;
}

int srt_repair(srt_ctx* c, const srt_params* p, const float* L_in, float* image_out)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !L_in || !image_out || L_in == image_out) return fail(c, SRT_ERR_INVALID, "repair needs distinct non-NULL buffers");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const size_t pixels = (size_t)c->camera.width * c->camera.height;
    int rc = ensure_frame(c, pixels);
    if (rc) return rc;
    SRT_CUDA(c, cudaMemcpyAsync(c->d_L, L_in, pixels * 4, cudaMemcpyHostToDevice, c->stream));
    rc = srt_repair_device(c, p, c->d_L, c->d_img);
    if (rc) return rc;
    if ((rc = copy_to_host(c, image_out, c->d_img, pixels * 4)) != 0) return rc;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (int erc = elapsed_ms(c, &c->timing.repair_ms, 2, 3)) return erc;
    c->timing.launches = 1;
    return SRT_OK;
}

// ---- whole frame --------------------------------------------------------------------
int srt_render_xray(srt_ctx* c, const srt_params* p, float* image_out)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !image_out) return fail(c, SRT_ERR_INVALID, "params / image_out is NULL");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const size_t pixels = (size_t)c->camera.width * c->camera.height;
    int rc = ensure_frame(c, pixels);
    if (rc) return rc;
    rc = render_device(c, p, 0, c->camera.height, c->d_L, nullptr, nullptr, nullptr);
    if (rc) return rc;
    rc = srt_repair_device(c, p, c->d_L, c->d_img);
    if (rc) return rc;
    if ((rc = copy_to_host(c, image_out, c->d_img, pixels * 4)) != 0) return rc;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (int erc = elapsed_ms(c, &c->timing.render_ms, 0, 1)) return erc;
    if (int erc = elapsed_ms(c, &c->timing.repair_ms, 2, 3)) return erc;
    c->timing.launches = 2;
    return SRT_OK;
}

// ---- sort-based X-ray (src/main.cxx:626-743) -------------------------------------------------
int srt_render_xray_sorted_device(srt_ctx* c, const srt_params* p, float* d_image, float* d_path_len, uint32_t* d_hit_count)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !d_image) return fail(c, SRT_ERR_INVALID, "params / image is NULL");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    return render_device(c, p, 0, c->camera.height, d_image, d_path_len, nullptr, d_hit_count, 0, 0, /*mode*/ 2);
}

int srt_render_xray_sorted(srt_ctx* c, const srt_params* p, float* image_out)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !image_out) return fail(c, SRT_ERR_INVALID, "params / image_out is NULL");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const size_t pixels = (size_t)c->camera.width * c->camera.height;
    int rc = ensure_frame(c, pixels);
    if (rc) return rc;
    rc = render_device(c, p, 0, c->camera.height, c->d_img, nullptr, nullptr, nullptr, 0, 0, /*mode*/ 2);
    if (rc) return rc;
    if ((rc = copy_to_host(c, image_out, c->d_img, pixels * 4)) != 0) return rc;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (int erc = elapsed_ms(c, &c->timing.render_ms, 0, 1)) return erc;
    c->timing.repair_ms = 0.0f;
    return SRT_OK;
}

// ---- shaded variant ------------------------------------------------------------------------
namespace {

// Build (or reuse) the light-space grid the shadow rays look their blockers up in: kernels.h ShadowGrid.
// Synchronous the first time a (scene, light) pair is seen (two small read-backs: the scene's bounding box and the
// number of list entries); cached afterwards.  Leaves grid.enabled = 0 (brute force) when the light is inside or
// close to the scene's bounding box, where a central projection onto one plane does not cover the scene.
int build_shadow_grid(srt_ctx* c, srt::ShadeArgs& a, uint32_t W, uint32_t H)
{
    srt::ShadowGrid& g = c->grid;
    const bool same_light = std::memcmp(c->grid_light, a.light_position, sizeof c->grid_light) == 0;
    if (c->grid_scene == c->scene_serial && same_light) { a.grid = g; return SRT_OK; }
    c->timing.shadow_grid_rebuilt = 1;
    const uint32_t total = (uint32_t)c->total_tris;
    g.enabled = 0; g.total_tris = total;
    a.grid = g;
    std::memcpy(c->grid_light, a.light_position, sizeof c->grid_light);
    c->grid_scene = c->scene_serial;
    if (c->force_shadow_grid < 0 || total == 0) return SRT_OK;
    if (!c->d_bbox) SRT_CUDA(c, cudaMalloc(&c->d_bbox, 6 * sizeof(uint32_t)));
    if (!g.n_global) SRT_CUDA(c, cudaMalloc(&g.n_global, sizeof(uint32_t)));
    if (c->bbox_scene != c->scene_serial) {
        const uint32_t init[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u };
        uint32_t keys[6];
        SRT_CUDA(c, cudaMemcpyAsync(c->d_bbox, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
        SRT_CUDA(c, srt::launch_scene_bbox(c->d_shade_meshes, (int)c->n_meshes, total, c->d_bbox, c->stream));
        SRT_CUDA(c, cudaMemcpyAsync(keys, c->d_bbox, sizeof keys, cudaMemcpyDeviceToHost, c->stream));
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 6; ++k) c->scene_bbox[k] = srt::ordered_key_to_float(keys[k]);
        c->bbox_scene = c->scene_serial;
        c->timing.launches += 1;
    }
    const float* bb = c->scene_bbox;
    const float* L = a.light_position;
    for (int k = 0; k < 3; ++k) if (!(bb[k] <= bb[3 + k])) return SRT_OK;          // empty or NaN box
    // axis: from the light to the centre of the box
    double w[3], len = 0.0;
    for (int k = 0; k < 3; ++k) { w[k] = 0.5 * ((double)bb[k] + bb[3 + k]) - L[k]; len += w[k] * w[k]; }
    len = std::sqrt(len);
    if (!(len > 0.0) || !std::isfinite(len)) return SRT_OK;
    for (int k = 0; k < 3; ++k) w[k] /= len;
    int least = 0;
    for (int k = 1; k < 3; ++k) if (std::fabs(w[k]) < std::fabs(w[least])) least = k;
    double e[3] = { 0, 0, 0 }, u[3], v[3];
    e[least] = 1.0;
    u[0] = w[1] * e[2] - w[2] * e[1]; u[1] = w[2] * e[0] - w[0] * e[2]; u[2] = w[0] * e[1] - w[1] * e[0];
    const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int k = 0; k < 3; ++k) u[k] /= ul;
    v[0] = w[1] * u[2] - w[2] * u[1]; v[1] = w[2] * u[0] - w[0] * u[2]; v[2] = w[0] * u[1] - w[1] * u[0];
    // projected extent of the box: all 8 corners must lie well in front of the light
    double amin = 1e300, amax = -1e300, bmin = 1e300, bmax = -1e300;
    for (int corner = 0; corner < 8; ++corner) {
        double q[3], qn = 0.0;
        for (int k = 0; k < 3; ++k) { q[k] = (double)bb[(corner >> k & 1) ? 3 + k : k] - L[k]; qn += std::fabs(q[k]); }
        const double depth = q[0] * w[0] + q[1] * w[1] + q[2] * w[2];
        if (!(depth > 0.05 * qn)) return SRT_OK;                                  // light inside / beside the box: brute force
        const double pa = (q[0] * u[0] + q[1] * u[1] + q[2] * u[2]) / depth, pb = (q[0] * v[0] + q[1] * v[1] + q[2] * v[2]) / depth;
        amin = std::min(amin, pa); amax = std::max(amax, pa); bmin = std::min(bmin, pb); bmax = std::max(bmax, pb);
    }
    const double ra = std::max(amax - amin, 1e-30), rb = std::max(bmax - bmin, 1e-30);
    uint32_t G = c->force_shadow_grid > 0 ? (uint32_t)c->force_shadow_grid
                                          : (uint32_t)std::max(std::sqrt((double)W * H) / 4.0, std::sqrt((double)total));
    G = std::min(std::max(G, 64u), 2048u);
    for (int k = 0; k < 3; ++k) { g.u[k] = (float)u[k]; g.v[k] = (float)v[k]; g.w[k] = (float)w[k]; }
    g.G = G;
    g.a0 = (float)(amin - 0.01 * ra); g.b0 = (float)(bmin - 0.01 * rb);           // 1 % apron around the box's projection
    g.inv_da = (float)(G / (1.02 * ra)); g.inv_db = (float)(G / (1.02 * rb));
    const size_t cells = (size_t)G * G;
    if (cells + 1 > c->grid_cells_cap) {
        cudaFree(g.cell_off); cudaFree(g.cell_cursor); g.cell_off = g.cell_cursor = nullptr; c->grid_cells_cap = 0;
        SRT_CUDA(c, cudaMalloc(&g.cell_off, (cells + 1) * sizeof(uint32_t)));
        SRT_CUDA(c, cudaMalloc(&g.cell_cursor, (cells + 1) * sizeof(uint32_t)));
        c->grid_cells_cap = cells + 1;
    }
    if (total > c->grid_global_cap) {
        cudaFree(g.global_tris); g.global_tris = nullptr; c->grid_global_cap = 0;
        SRT_CUDA(c, cudaMalloc(&g.global_tris, (size_t)total * sizeof(uint32_t)));
        c->grid_global_cap = total;
    }
    a.grid = g;
    SRT_CUDA(c, cudaMemsetAsync(g.cell_off, 0, (cells + 1) * sizeof(uint32_t), c->stream));
    SRT_CUDA(c, srt::launch_shadow_bin(a, 0, c->stream));
    SRT_CUDA(c, srt::launch_exclusive_scan(g.cell_off, (uint32_t)(cells + 1), c->stream));
    uint32_t entries = 0;
    SRT_CUDA(c, cudaMemcpyAsync(&entries, g.cell_off + cells, sizeof entries, cudaMemcpyDeviceToHost, c->stream));
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if ((size_t)entries > c->grid_tris_cap) {
        cudaFree(g.cell_tris); g.cell_tris = nullptr; c->grid_tris_cap = 0;
        const size_t cap = (size_t)entries + entries / 8 + 1024;
        SRT_CUDA(c, cudaMalloc(&g.cell_tris, cap * sizeof(uint32_t)));
        c->grid_tris_cap = cap;
    }
    a.grid = g;
    SRT_CUDA(c, cudaMemcpyAsync(g.cell_cursor, g.cell_off, cells * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    SRT_CUDA(c, cudaMemsetAsync(g.n_global, 0, sizeof(uint32_t), c->stream));
    SRT_CUDA(c, srt::launch_shadow_bin(a, 1, c->stream));
    g.enabled = 1;
    a.grid = g;
    c->timing.launches += 3;
    c->timing.shadow_grid_cells = G;
    c->timing.shadow_grid_entries = entries;
    return SRT_OK;
}

} // namespace

int srt_render_shaded_device(srt_ctx* c, const srt_params* p, const srt_light* light, const uint8_t background_rgb[3],
                             uint8_t* d_rgb, uint8_t* d_shadow)
{
    if (!c) return SRT_ERR_INVALID;
    if (!p || !light || !background_rgb || !d_rgb) return fail(c, SRT_ERR_INVALID, "params / light / background / rgb is NULL");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const uint32_t W = c->camera.width, H = c->camera.height;
    int rc = ensure_frame(c, (size_t)W * H);
    if (rc) return rc;
    // pass 1: closest hit over all meshes (z-buffer + triangle index per pixel)
    rc = render_device(c, p, 0, H, nullptr, c->d_dist, c->d_ssum, nullptr, 0, 0, /*mode*/ 1);
    if (rc) return rc;
    // pass 2: shading, shadow rays, texture
    srt::ShadeArgs a{};
    const srt_camera& cam = c->camera;
    a.cam.ox = cam.origin[0]; a.cam.oy = cam.origin[1]; a.cam.oz = cam.origin[2];
    a.cam.dx = cam.detector[0]; a.cam.dy = cam.detector[1]; a.cam.dz = cam.detector[2];
    a.cam.ux = cam.up[0]; a.cam.uy = cam.up[1]; a.cam.uz = cam.up[2];
    a.cam.rx = cam.right[0]; a.cam.ry = cam.right[1]; a.cam.rz = cam.right[2];
    a.cam.u_off = c->d_u_off; a.cam.v_off = c->d_v_off;
    a.cam.xc = (float)(0.5 - (double)W / 2.0); a.cam.yc = (float)(0.5 - (double)H / 2.0);
    a.cam.W = W; a.cam.H = H;
    a.meshes = c->d_shade_meshes; a.n_meshes = (int)c->n_meshes;
    for (int k = 0; k < 3; ++k) { a.light_colour[k] = light->colour[k]; a.light_position[k] = light->position[k]; a.background[k] = background_rgb[k]; }
    if (!d_shadow) {                                        // kept for srt_debug_shaded_buffers
        if ((size_t)W * H > c->shadow_cap) {
            SRT_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->d_shadow); c->d_shadow = nullptr; c->shadow_cap = 0;
            SRT_CUDA(c, cudaMalloc(&c->d_shadow, (size_t)W * H)); c->shadow_cap = (size_t)W * H;
        }
        d_shadow = c->d_shadow;
    }
    a.z = c->d_dist; a.tri_id = c->d_ssum; a.rgb = d_rgb; a.shadow = d_shadow;
    a.row_begin = 0; a.row_end = H;
    c->timing.shadow_grid_rebuilt = 0;
    SRT_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    rc = build_shadow_grid(c, a, W, H);
    if (rc) return rc;
    c->timing.shadow_grid_cells = a.grid.enabled ? a.grid.G : 0;
    SRT_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    SRT_CUDA(c, srt::launch_shade(a, c->stream));
    SRT_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    c->timing.launches += 1;
    return SRT_OK;
}

int srt_render_shaded(srt_ctx* c, const srt_params* p, const srt_light* light, const uint8_t background_rgb[3], uint8_t* rgb_out)
{
    if (!c) return SRT_ERR_INVALID;
    if (!rgb_out) return fail(c, SRT_ERR_INVALID, "rgb_out is NULL");
    if (!c->have_camera) return fail(c, SRT_ERR_INVALID, "srt_set_camera has not been called");
    SRT_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->camera.width * c->camera.height * 3;
    if (bytes > c->rgb_cap) {
        SRT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_rgb); c->d_rgb = nullptr;
        SRT_CUDA(c, cudaMalloc(&c->d_rgb, bytes)); c->rgb_cap = bytes;
    }
    int rc = srt_render_shaded_device(c, p, light, background_rgb, c->d_rgb, nullptr);
    if (rc) return rc;
    if ((rc = copy_to_host(c, rgb_out, c->d_rgb, bytes)) != 0) return rc;
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (int erc = elapsed_ms(c, &c->timing.render_ms, 0, 1)) return erc;
    if (int erc = elapsed_ms(c, &c->timing.repair_ms, 2, 3)) return erc;     // shade pass time
    if (int erc = elapsed_ms(c, &c->timing.shadow_grid_ms, 6, 2)) return erc;
    c->timing_pending = false;
    return SRT_OK;
}

// diagnostic: z-buffer / triangle index / shadow flag of the last shaded render (host copies; any may be NULL)
int srt_debug_shaded_buffers(srt_ctx* c, float* z_out, int32_t* tri_id_out, uint8_t* shadow_out)
{
    if (!c || !c->have_camera) return SRT_ERR_INVALID;
    const size_t n = (size_t)c->camera.width * c->camera.height;
    if (n > c->frame_cap) return fail(c, SRT_ERR_INVALID, "no shaded render yet");
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (z_out) SRT_CUDA(c, cudaMemcpy(z_out, c->d_dist, n * 4, cudaMemcpyDeviceToHost));
    if (tri_id_out) SRT_CUDA(c, cudaMemcpy(tri_id_out, c->d_ssum, n * 4, cudaMemcpyDeviceToHost));
    if (shadow_out) {
        if (n > c->shadow_cap) return fail(c, SRT_ERR_INVALID, "the last shaded render wrote its shadow flags to a caller's buffer");
        SRT_CUDA(c, cudaMemcpy(shadow_out, c->d_shadow, n, cudaMemcpyDeviceToHost));
    }
    return SRT_OK;
}

int srt_last_timing(srt_ctx* c, srt_timing* out)
{
    if (!c || !out) return SRT_ERR_INVALID;
    if (c->timing_pending) {
        SRT_CUDA(c, cudaSetDevice(c->device));
        if (c->profiling) SRT_CUDA(c, cudaEventSynchronize(c->ev[1]));
        if (int erc = elapsed_ms(c, &c->timing.render_ms, 0, 1)) return erc;
        if (int erc = elapsed_ms(c, &c->timing.main_kernel_ms, 4, 5)) return erc;
        if (c->counters_pending && c->d_counters) {
            unsigned long long cnts[3] = { 0, 0, 0 };
            SRT_CUDA(c, cudaMemcpyAsync(cnts, c->d_counters, sizeof cnts, cudaMemcpyDeviceToHost, c->stream));
            SRT_CUDA(c, cudaStreamSynchronize(c->stream));
            c->timing.exact_tests = cnts[0]; c->timing.filter_violations = cnts[1];
            std::memcpy(&c->timing.min_strip_slack, &cnts[2], sizeof(float));
            std::memcpy(&c->timing.min_box_slack, (const char*)&cnts[2] + sizeof(float), sizeof(float));
        }
        c->timing_pending = false; c->counters_pending = false;
    }
    *out = c->timing;
    return SRT_OK;
}

// diagnostic: download the device-side packed arrays of one mesh (packs the scene first if needed, for the
// default box shape unless a render has packed another one)
int srt_debug_download_packed(srt_ctx* c, int mesh, float* exact16, float* filt12, float* filt2_8)
{
    if (!c) return SRT_ERR_INVALID;
    SRT_CUDA(c, cudaSetDevice(c->device));
    int rc = c->packed ? SRT_OK : pack_scene(c);
    if (rc) return rc;
    if (mesh < 0 || (size_t)mesh >= c->n_meshes) return fail(c, SRT_ERR_INVALID, "mesh index %d out of range", mesh);
    const MeshBuf& mb = c->meshes[mesh];
    SRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (exact16 && mb.n) SRT_CUDA(c, cudaMemcpy(exact16, mb.exact, (size_t)mb.n * sizeof(srt::ExactTri), cudaMemcpyDeviceToHost));
    if (filt12 && mb.n_padded) SRT_CUDA(c, cudaMemcpy(filt12, mb.filt0(), (size_t)mb.n_padded * sizeof(srt::FilterTri), cudaMemcpyDeviceToHost));
    if (filt2_8 && mb.n_padded) SRT_CUDA(c, cudaMemcpy(filt2_8, mb.filt1(), (size_t)mb.n_padded * sizeof(srt::FilterTri2), cudaMemcpyDeviceToHost));
    return SRT_OK;
}

// diagnostic: how meshes are packed -- padding granularity (triangles) and the default box of the filtered kernel
int srt_debug_pack_info(uint32_t* pad_to, uint32_t* box_w, uint32_t* box_h)
{
    if (pad_to) *pad_to = srt::kPadTris;
    if (box_w) *box_w = SRT_BOX_PX;
    if (box_h) *box_h = SRT_BOX_PY;
    return SRT_OK;
}

// diagnostic (host only, no GPU): run the pre-pack for one mesh and return the packed arrays.
// exact16: 16 floats per triangle; filt12: 12, filt2_8: 8 floats per padded triangle.  box_w x box_h: the box the
// level-0 bounds FilterTri::R / R2 (filt12[7], [11]) are packed for (0 = the default).
int srt_debug_prepack(const float* vertices, uint32_t n, const srt_camera* cam, uint32_t box_w, uint32_t box_h,
                      float* exact16, float* filt12, float* filt2_8, uint32_t* n_padded, uint32_t* n_unfilterable)
{
    if (!vertices || !cam) return SRT_ERR_INVALID;
    if (box_w == 0) { box_w = SRT_BOX_PX; box_h = SRT_BOX_PY; }
    if (box_h == 0) box_h = 1;
    srt::FrameGeom g = srt::make_frame_geom(cam->origin, cam->detector, cam->up, cam->right, cam->pixel_spacing[0],
                                            cam->pixel_spacing[1], cam->width, cam->height, box_w, box_h);
    srt::PackedMesh pm;
    srt::prepack_mesh(vertices, nullptr, n, cam->origin, g, srt::kPadTris, pm);
    if (exact16) std::memcpy(exact16, pm.exact.data(), pm.exact.size() * sizeof(srt::ExactTri));
    if (filt12) std::memcpy(filt12, pm.filt.data(), pm.filt.size() * sizeof(srt::FilterTri));
    if (filt2_8) std::memcpy(filt2_8, pm.filt2.data(), pm.filt2.size() * sizeof(srt::FilterTri2));
    if (n_padded) *n_padded = pm.n_padded;
    if (n_unfilterable) *n_unfilterable = pm.unfilterable;
    return SRT_OK;
}

// diagnostic: measured FP32 FMA peak of this GPU in TFLOP/s (2 flop per FFMA), best of `repeats` launches
int srt_debug_ffma_peak(srt_ctx* c, int repeats, float* tflops_out)
{
    if (!c || !tflops_out) return SRT_ERR_INVALID;
    SRT_CUDA(c, cudaSetDevice(c->device));
    cudaDeviceProp prop{};
    SRT_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 14;
    float* scratch = nullptr;
    SRT_CUDA(c, cudaMalloc(&scratch, (size_t)blocks * 256 * sizeof(float)));
    float best = 0.0f;
    for (int r = 0; r < (repeats < 1 ? 1 : repeats) + 1; ++r) {       // first launch is the warm-up
        SRT_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
        SRT_CUDA(c, srt::launch_ffma_peak(scratch, blocks, iters, c->stream));
        SRT_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
        SRT_CUDA(c, cudaEventSynchronize(c->ev[3]));
        float ms = 0.0f;
        SRT_CUDA(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
        const double flop = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
        const float tf = (float)(flop / (ms * 1e-3) / 1e12);
        if (r > 0 && tf > best) best = tf;
    }
    cudaFree(scratch);
    *tflops_out = best;
    return SRT_OK;
}

// diagnostic: force the number of triangle-range splits of the filtered kernel (0 = heuristic)
int srt_debug_force_scan_cols(srt_ctx* c, int cols)
{
    if (!c || cols < 0) return SRT_ERR_INVALID;
    c->force_scan_cols = cols;
    return SRT_OK;
}

// diagnostic: cap the culled kernel's per-bin lists and the filtered path's candidate lists (0 = no cap) so that tests
// can exercise the overflow fallbacks
int srt_debug_force_list_caps(srt_ctx* c, uint32_t bin_cap, uint32_t cand_cap, uint32_t pix_cap)
{
    if (!c) return SRT_ERR_INVALID;
    c->force_bin_cap = bin_cap; c->force_cand_cap = cand_cap; c->force_pix_cap = pix_cap;
    return SRT_OK;
}

// diagnostic: cells per side of the shadow rays' light-space grid (0 = heuristic, < 0 = no grid: every shadow ray
// tests every triangle, the reference's loop)
int srt_debug_force_shadow_grid(srt_ctx* c, int cells_per_side)
{
    if (!c) return SRT_ERR_INVALID;
    c->force_shadow_grid = cells_per_side;
    c->grid_scene = 0;
    return SRT_OK;
}

// test/diagnostic hook (not part of the reference-facing surface): count exact-path tests
int srt_debug_count_exact(srt_ctx* c, int enable)
{
    if (!c) return SRT_ERR_INVALID;
    c->count_exact = enable != 0;
    return SRT_OK;
}


// ---- one process, N GPUs -------------------------------------------------------------------------
} // extern "C"

struct srt_multi {
    int n = 0;
    std::vector<int> dev;
    std::vector<srt_ctx*> ctx;
    std::vector<float*> L[2];            // two frame sets used alternately (one cross-device barrier per step suffices)
    std::vector<float*> img;
    std::vector<cudaEvent_t> ev_start, ev_rendered, ev_repaired, ev_end;
    size_t frame_cap = 0;
    unsigned step = 0;
    bool have_camera = false; uint32_t W = 0, H = 0;
    std::string error;
};

namespace {

int mfail(srt_multi* m, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (m) m->error = buf; else g_create_error = buf;
    return code;
}

#define SRT_MCUDA(m, call)                                                                       \
    do { cudaError_t e__ = (call);                                                               \
         if (e__ != cudaSuccess) return mfail((m), SRT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); } while (0)

void multi_free_frames(srt_multi* m)
{
    for (int i = 0; i < m->n; ++i) {
        cudaSetDevice(m->dev[i]);
        for (int s = 0; s < 2; ++s) { if (i < (int)m->L[s].size()) cudaFree(m->L[s][i]); }
        if (i < (int)m->img.size()) cudaFree(m->img[i]);
    }
    m->L[0].assign(m->n, nullptr); m->L[1].assign(m->n, nullptr); m->img.assign(m->n, nullptr);
    m->frame_cap = 0;
}

int multi_ensure_frames(srt_multi* m, size_t pixels)
{
    if (pixels <= m->frame_cap) return SRT_OK;
    int rc = srt_multi_synchronize(m);
    if (rc) return rc;
    multi_free_frames(m);
    for (int i = 0; i < m->n; ++i) {
        SRT_MCUDA(m, cudaSetDevice(m->dev[i]));
        for (int s = 0; s < 2; ++s) SRT_MCUDA(m, cudaMalloc(&m->L[s][i], pixels * sizeof(float)));
        SRT_MCUDA(m, cudaMalloc(&m->img[i], pixels * sizeof(float)));
    }
    m->frame_cap = pixels;
    return SRT_OK;
}

} // namespace

extern "C" {

int srt_multi_create(srt_multi** out, const int* device_ids, int n_devices)
{
    if (!out) return mfail(nullptr, SRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 8) return mfail(nullptr, SRT_ERR_INVALID, "1 to 8 device ids expected");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return mfail(nullptr, SRT_ERR_INVALID, "device %d listed twice", device_ids[i]);
    srt_multi* m = new (std::nothrow) srt_multi();
    if (!m) return mfail(nullptr, SRT_ERR_NOMEM, "out of host memory");
    m->n = n_devices;
    m->dev.assign(device_ids, device_ids + n_devices);
    for (int i = 0; i < n_devices; ++i) {
        srt_ctx* c = nullptr;
        const int rc = srt_create(&c, device_ids[i]);
        if (rc) { srt_multi_destroy(m); return rc; }              // message already in the create-error slot
        m->ctx.push_back(c);
    }
    // every device maps every other device's memory: the render kernels store into the peers' frames
    for (int i = 0; i < n_devices; ++i) {
        cudaSetDevice(device_ids[i]);
        for (int j = 0; j < n_devices; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]);
            if (!can) {
                srt_multi_destroy(m);
                return mfail(nullptr, SRT_ERR_NO_DEVICE, "device %d cannot access device %d's memory (peer access is required)",
                             device_ids[i], device_ids[j]);
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                srt_multi_destroy(m);
                return mfail(nullptr, SRT_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device_ids[i], device_ids[j], cudaGetErrorString(e));
            }
            cudaGetLastError();                                    // clear "already enabled"
        }
    }
    m->ev_start.assign(n_devices, nullptr); m->ev_rendered.assign(n_devices, nullptr); m->ev_repaired.assign(n_devices, nullptr);
    m->ev_end.assign(n_devices, nullptr);
    m->L[0].assign(n_devices, nullptr); m->L[1].assign(n_devices, nullptr); m->img.assign(n_devices, nullptr);
    for (int i = 0; i < n_devices; ++i) {
        cudaSetDevice(device_ids[i]);
        if (cudaEventCreate(&m->ev_start[i]) != cudaSuccess || cudaEventCreate(&m->ev_end[i]) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_rendered[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_repaired[i], cudaEventDisableTiming) != cudaSuccess) {
            srt_multi_destroy(m);
            return mfail(nullptr, SRT_ERR_CUDA, "event creation failed on device %d", device_ids[i]);
        }
    }
    *out = m;
    return SRT_OK;
}

void srt_multi_destroy(srt_multi* m)
{
    if (!m) return;
    for (size_t i = 0; i < m->ctx.size(); ++i) { cudaSetDevice(m->dev[i]); cudaStreamSynchronize(m->ctx[i]->stream); }
    if (!m->img.empty()) multi_free_frames(m);
    for (size_t i = 0; i < m->ev_start.size(); ++i) {
        cudaSetDevice(m->dev[i]);
        if (m->ev_start[i]) cudaEventDestroy(m->ev_start[i]);
        if (m->ev_rendered[i]) cudaEventDestroy(m->ev_rendered[i]);
        if (i < m->ev_repaired.size() && m->ev_repaired[i]) cudaEventDestroy(m->ev_repaired[i]);
        if (m->ev_end[i]) cudaEventDestroy(m->ev_end[i]);
    }
    for (srt_ctx* c : m->ctx) srt_destroy(c);
    delete m;
}

const char* srt_multi_last_error(const srt_multi* m) { return m ? m->error.c_str() : g_create_error.c_str(); }

int srt_multi_device_count(const srt_multi* m) { return m ? m->n : 0; }

int srt_multi_set_meshes(srt_multi* m, int n_meshes, const srt_mesh_desc* meshes)
{
    if (!m) return SRT_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        const int rc = srt_set_meshes(m->ctx[i], n_meshes, meshes);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[i], m->ctx[i]->error.c_str());
    }
    return SRT_OK;
}

int srt_multi_set_camera(srt_multi* m, const srt_camera* cam)
{
    if (!m) return SRT_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        const int rc = srt_set_camera(m->ctx[i], cam);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[i], m->ctx[i]->error.c_str());
    }
    m->have_camera = true; m->W = cam->width; m->H = cam->height;
    return SRT_OK;
}

int srt_multi_synchronize(srt_multi* m)
{
    if (!m) return SRT_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        SRT_MCUDA(m, cudaSetDevice(m->dev[i]));
        SRT_MCUDA(m, cudaStreamSynchronize(m->ctx[i]->stream));
    }
    return SRT_OK;
}

int srt_multi_render_xray_device(srt_multi* m, const srt_params* params)
{
    if (!m) return SRT_ERR_INVALID;
    if (!params) return mfail(m, SRT_ERR_INVALID, "params is NULL");
    if (!m->have_camera) return mfail(m, SRT_ERR_INVALID, "srt_multi_set_camera has not been called");
    int rc = multi_ensure_frames(m, (size_t)m->W * m->H);
    if (rc) return rc;
    const unsigned set = m->step & 1u;
    // 1. every device renders its blocks: L rows into its own frame, the rows the neighbouring devices' repair stencils
    //    reach (the first / last repair_depth rows of a block) also into those devices' frames over NVLink
    for (int i = 0; i < m->n; ++i) {
        srt_ctx* c = m->ctx[i];
        SRT_MCUDA(m, cudaSetDevice(m->dev[i]));
        SRT_MCUDA(m, cudaEventRecord(m->ev_start[i], c->stream));
        float* peers[7]; uint32_t np = 0;
        for (int j = 0; j < m->n; ++j) if (j != i) peers[np++] = m->L[set][j];      // rank order, own skipped
        rc = render_device(c, params, 0, m->H, m->L[set][i], nullptr, nullptr, nullptr, (uint32_t)m->n, (uint32_t)i, 0,
                           np ? peers : nullptr, np, /*image_positions*/ true, /*multicast*/ false, /*halo*/ np > 0);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[i], c->error.c_str());
        SRT_MCUDA(m, cudaEventRecord(m->ev_rendered[i], c->stream));
    }
    // 2. cross-device barrier (events), then the sharded repair pass: every device repairs the pixels of its own blocks and
    //    stores them into the image of every device
    for (int i = 0; i < m->n; ++i) {
        srt_ctx* c = m->ctx[i];
        SRT_MCUDA(m, cudaSetDevice(m->dev[i]));
        for (int j = 0; j < m->n; ++j) if (j != i) SRT_MCUDA(m, cudaStreamWaitEvent(c->stream, m->ev_rendered[j], 0));
        float* images[8]; uint32_t ni = 0;
        images[ni++] = m->img[i];
        for (int j = 0; j < m->n; ++j) if (j != i) images[ni++] = m->img[j];
        rc = srt_repair_blocks_device(c, params, (uint32_t)m->n, (uint32_t)i, m->L[set][i], ni, images);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[i], c->error.c_str());
        SRT_MCUDA(m, cudaEventRecord(m->ev_repaired[i], c->stream));
    }
    // 3. second barrier: every device's image is complete when all repair passes have run
    for (int i = 0; i < m->n; ++i) {
        srt_ctx* c = m->ctx[i];
        SRT_MCUDA(m, cudaSetDevice(m->dev[i]));
        for (int j = 0; j < m->n; ++j) if (j != i) SRT_MCUDA(m, cudaStreamWaitEvent(c->stream, m->ev_repaired[j], 0));
        SRT_MCUDA(m, cudaEventRecord(m->ev_end[i], c->stream));
    }
    ++m->step;
    return SRT_OK;
}

int srt_multi_frame(srt_multi* m, int index, float** d_L, float** d_image)
{
    if (!m || index < 0 || index >= m->n) return SRT_ERR_INVALID;
    if (m->step == 0 || m->frame_cap == 0) return mfail(m, SRT_ERR_INVALID, "no frame has been rendered yet");
    const unsigned set = (m->step - 1) & 1u;
    if (d_L) *d_L = m->L[set][index];
    if (d_image) *d_image = m->img[index];
    return SRT_OK;
}

int srt_multi_render_xray(srt_multi* m, const srt_params* params, float* image_out)
{
    if (!m) return SRT_ERR_INVALID;
    if (!image_out) return mfail(m, SRT_ERR_INVALID, "image_out is NULL");
    int rc = srt_multi_render_xray_device(m, params);
    if (rc) return rc;
    SRT_MCUDA(m, cudaSetDevice(m->dev[0]));
    if ((rc = copy_to_host(m->ctx[0], image_out, m->img[0], (size_t)m->W * m->H * sizeof(float))) != 0)
        return mfail(m, rc, "%s", m->ctx[0]->error.c_str());
    return srt_multi_synchronize(m);
}

int srt_multi_last_timing(srt_multi* m, float* render_ms_per_device, float* step_ms)
{
    if (!m) return SRT_ERR_INVALID;
    if (m->step == 0) return mfail(m, SRT_ERR_INVALID, "no frame has been rendered yet");
    int rc = srt_multi_synchronize(m);
    if (rc) return rc;
    float worst = 0.0f;
    for (int i = 0; i < m->n; ++i) {
        SRT_MCUDA(m, cudaSetDevice(m->dev[i]));
        srt_timing t{};
        rc = srt_last_timing(m->ctx[i], &t);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[i], m->ctx[i]->error.c_str());
        if (render_ms_per_device) render_ms_per_device[i] = t.render_ms;
        float ms = 0.0f;
        SRT_MCUDA(m, cudaEventElapsedTime(&ms, m->ev_start[i], m->ev_end[i]));
        worst = std::max(worst, ms);
    }
    if (step_ms) *step_ms = worst;
    return SRT_OK;
}

} // extern "C"
