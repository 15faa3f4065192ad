/*
 * filter_audit.c -- CPU audit of the product's conservative reject filter.
 *
 * TEST INFRASTRUCTURE ONLY (see srt_oracle.c).  Takes the packed per-triangle
 * filter constants produced by the product's host pre-pack (srt_debug_prepack)
 * and, for every (pixel, triangle) pair of a frame, evaluates the filter with
 * the same fused-multiply-add sequence the CUDA kernels use (fmaf == one
 * rounding, like FFMA) next to the oracle's Ray::intersect restatement.
 * A "violation" is a pair the filter rejects although the oracle reports a hit;
 * the filter is only valid if there are none.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float detector[3], origin[3], up[3], right[3];
    float lower[3], upper[3], range[3];
    float light_position[3], light_direction[3];
} oracle_camera;

extern void oracle_make_ray(const oracle_camera* cam, uint32_t W, uint32_t H, uint32_t row, uint32_t col,
                            float dirA[3], float dirB[3]);
extern int oracle_intersect(const float o[3], const float d[3], const float tri9[9], float* t, int* stage);

typedef struct {
    const float* tri; uint32_t n; const oracle_camera* cam; uint32_t W, H;
    const float* filt; const float* filt2; uint32_t box_w, box_h;
    uint64_t p0, p1;
    /* out: [0] pairs, [1] pairs whose box survives the level-0 box gate of the filtered kernel (|Y(c)| - |D(c)| > R for
     *      either of its two strips, at the centre of the aligned box_w x box_h box, clamped to the frame; kernels.cu scan_tile), [2] survive the
     *      per-pixel level-0 strip as well, [3] survive all three per-pixel strips, [4] oracle hits, [5] violations:
     *      hits that the level-0 box gate or a per-pixel strip rejects, [6] hits that the three-strip test of the
     *      filtered kernel's box rejects (kernels.cu drain_q1), [7] hits inside a box that the box test of the culled
     *      kernel rejects (bin 256 x 128, CTA tile 64 x 16, warp tile 32 x 8, run 8 x 1 of a full-frame launch) */
    uint64_t c[8];
} audit_job;

/* kernels.cu box_may_hit, same operation order (fmaf == FFMA; the plain products and the final scaling are separate
 * roundings there as well) */
__attribute__((target("fma"))) static int box_may_hit(const float* f, const float* g, float xc, float yc, float hx, float hy)
{
    const float infl = 1.0000038f;
    const float Dc = fabsf(__builtin_fmaf(f[3], xc, __builtin_fmaf(f[4], yc, f[5])));
    const float rD = __builtin_fmaf(fabsf(f[3]), hx, fabsf(f[4]) * hy);
    const float Y1 = fabsf(__builtin_fmaf(f[0], xc, __builtin_fmaf(f[1], yc, f[2])));
    const float Y2 = fabsf(__builtin_fmaf(g[0], xc, __builtin_fmaf(g[1], yc, g[2])));
    const float Y3 = fabsf(__builtin_fmaf(g[4], xc, __builtin_fmaf(g[5], yc, g[6])));
    const float s1 = __builtin_fmaf(fabsf(f[0]), hx, __builtin_fmaf(fabsf(f[1]), hy, __builtin_fmaf(2.0f, f[6], rD))) * infl;
    const float s2 = __builtin_fmaf(fabsf(g[0]), hx, __builtin_fmaf(fabsf(g[1]), hy, __builtin_fmaf(2.0f, g[3], rD))) * infl;
    const float s3 = __builtin_fmaf(fabsf(g[4]), hx, __builtin_fmaf(fabsf(g[5]), hy, __builtin_fmaf(2.0f, g[7], rD))) * infl;
    return !(Y1 - Dc > s1) && !(Y2 - Dc > s2) && !(Y3 - Dc > s3);
}

/* box [c0, c1] x [r0, r1] of pixel centres, clamped to the frame like the kernels clamp theirs */
__attribute__((target("fma"))) static int box_level(const float* f, const float* g, uint32_t W, uint32_t H, float xc0, float yc0,
                                                   uint32_t col, uint32_t row, uint32_t bw, uint32_t bh)
{
    uint32_t c0 = col - col % bw, c1 = c0 + bw, r0 = row - row % bh, r1 = r0 + bh;
    c1 = (c1 < W ? c1 : W) - 1; r1 = (r1 < H ? r1 : H) - 1;
    return box_may_hit(f, g, 0.5f * (float)(c0 + c1) + xc0, 0.5f * (float)(r0 + r1) + yc0, 0.5f * (float)(c1 - c0), 0.5f * (float)(r1 - r0));
}

__attribute__((target("fma"))) static void* audit_worker(void* arg)
{
    audit_job* j = (audit_job*)arg;
    const float xc = (float)(0.5 - (double)j->W / 2.0), yc = (float)(0.5 - (double)j->H / 2.0);
    for (uint64_t p = j->p0; p < j->p1; ++p) {
        uint32_t row = (uint32_t)(p / j->W), col = (uint32_t)(p % j->W);
        float dA[3], dB[3];
        oracle_make_ray(j->cam, j->W, j->H, row, col, dA, dB);
        const float x = (float)col + xc, y = (float)row + yc;
        /* the filtered kernel's box of this pixel: aligned box_w x box_h, clamped to the frame */
        uint32_t c0 = col - col % j->box_w, c1 = c0 + j->box_w, r0 = row - row % j->box_h, r1 = r0 + j->box_h;
        c1 = (c1 < j->W ? c1 : j->W) - 1; r1 = (r1 < j->H ? r1 : j->H) - 1;
        const float bxc = 0.5f * (float)(c0 + c1) + xc, byc = 0.5f * (float)(r0 + r1) + yc;
        for (uint32_t k = 0; k < j->n; ++k) {
            const float* f = j->filt + 12 * (size_t)k;
            const float* g = j->filt2 + 8 * (size_t)k;
            /* kernel: Y = fma(bx, x, fma(gx, y, ax)); D = fma(bd, x, fma(gd, y, ad)) */
            const float X1 = __builtin_fmaf(f[0], x, __builtin_fmaf(f[1], y, f[2]));
            const float Dd = __builtin_fmaf(f[3], x, __builtin_fmaf(f[4], y, f[5]));
            const float X2 = __builtin_fmaf(g[0], x, __builtin_fmaf(g[1], y, g[2]));
            const float X3 = __builtin_fmaf(g[4], x, __builtin_fmaf(g[5], y, g[6]));
            /* level-0 box gate (scan_tile): row part first, then the column part, |Y| - |D| against the packed R */
            const float Yc = __builtin_fmaf(f[0], bxc, __builtin_fmaf(f[1], byc, f[2]));
            const float Dc = __builtin_fmaf(f[3], bxc, __builtin_fmaf(f[4], byc, f[5]));
            const float Yc2 = __builtin_fmaf(f[8], bxc, __builtin_fmaf(f[9], byc, f[10]));
            const int s0t = !((fabsf(Yc) - fabsf(Dc) > f[7]) || (fabsf(Yc2) - fabsf(Dc) > f[11]));
            const int s0 = s0t && !(fabsf(X1) > fabsf(Dd) + f[6]);
            /* level 1: the two remaining strips */
            const int rej1 = (fabsf(X2) > fabsf(Dd) + g[3]) || (fabsf(X3) > fabsf(Dd) + g[7]);
            const int s1 = s0 && !rej1;
            float t; int st;
            const int hit = oracle_intersect(j->cam->origin, dB, j->tri + 9 * (size_t)k, &t, &st);
            j->c[0]++; j->c[1] += s0t; j->c[2] += s0; j->c[3] += s1; j->c[4] += hit;
            if (hit && !s1) j->c[5]++;
            if (hit) {
                /* filtered kernel, queue stage (drain_q1): three-strip test of the same box */
                if (!box_may_hit(f, g, bxc, byc, 0.5f * (float)(c1 - c0), 0.5f * (float)(r1 - r0))) j->c[6]++;
                /* culled kernel: every box that contains this pixel must let the triangle through */
                if (!box_level(f, g, j->W, j->H, xc, yc, col, row, 256, 128) || !box_level(f, g, j->W, j->H, xc, yc, col, row, 64, 16) ||
                    !box_level(f, g, j->W, j->H, xc, yc, col, row, 32, 8) || !box_level(f, g, j->W, j->H, xc, yc, col, row, 8, 1))
                    j->c[7]++;
            }
        }
    }
    return 0;
}

int oracle_filter_audit(const float* tri, uint32_t n, const oracle_camera* cam, uint32_t W, uint32_t H,
                        const float* filt, const float* filt2, uint32_t box_w, uint32_t box_h,
                        uint64_t pixel_begin, uint64_t pixel_end, int n_threads, uint64_t counts[8])
{
    if (!__builtin_cpu_supports("fma")) return -1;
    if (n_threads < 1) n_threads = 1;
    audit_job* jobs = (audit_job*)calloc((size_t)n_threads, sizeof(audit_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t total = pixel_end - pixel_begin;
    for (int i = 0; i < n_threads; ++i) {
        audit_job* j = &jobs[i];
        j->tri = tri; j->n = n; j->cam = cam; j->W = W; j->H = H; j->filt = filt; j->filt2 = filt2;
        j->box_w = box_w ? box_w : 1; j->box_h = box_h ? box_h : 1;
        j->p0 = pixel_begin + total * (uint64_t)i / (uint64_t)n_threads;
        j->p1 = pixel_begin + total * (uint64_t)(i + 1) / (uint64_t)n_threads;
        pthread_create(&tid[i], NULL, audit_worker, j);
    }
    memset(counts, 0, 8 * sizeof(uint64_t));
    for (int i = 0; i < n_threads; ++i) {
        pthread_join(tid[i], NULL);
        for (int k = 0; k < 8; ++k) counts[k] += jobs[i].c[k];
    }
    free(jobs); free(tid);
    return 0;
}
