// pipe_probe.cu -- issue-rate probe for the instructions the level-0 filter is built from (development aid).
// Prints warp-instructions per clock per SM for FFMA, HFMA2, HADD2, packed/scalar 3-input max and some mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CHAINS 16
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, const float* cf)
{
    float r[CHAINS]; __half2 h[CHAINS];
    const float a = cf[0], b = cf[1];
    const __half2 ha = __floats2half2_rn(cf[2], cf[3]), hb = __floats2half2_rn(cf[1], cf[0]);
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { r[k] = (float)(threadIdx.x + k); h[k] = __floats2half2_rn(0.001f * (threadIdx.x + k), 0.5f); }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) {
                if (MODE == 0) r[k] = fmaf(r[k], a, b);                                   // FFMA
                if (MODE == 1) h[k] = __hfma2(h[k], ha, hb);                              // HFMA2
                if (MODE == 2) h[k] = __hadd2(h[k], ha);                                  // HADD2
                if (MODE == 3) h[k] = __hmax2(h[k], __hmax2(h[(k + 1) % CHAINS], ha));    // VHMNMX (3-input)
                if (MODE == 4) asm("max.f32 %0, %0, %1, %2;" : "+f"(r[k]) : "f"(r[(k + 1) % CHAINS]), "f"(a));   // FMNMX3
                if (MODE == 5) { if (k & 1) r[k] = fmaf(r[k], a, b); else h[k] = __hfma2(h[k], ha, hb); }      // FFMA : HFMA2 1:1
                if (MODE == 6) { if (k & 1) h[k] = __hmax2(h[k], __hmax2(h[(k + 2) % CHAINS], ha)); else h[k] = __hfma2(h[k], ha, hb); }   // HFMA2 : VHMNMX 1:1
                if (MODE == 7) { if (k & 1) r[k] = fmaf(r[k], a, b); else asm("max.f32 %0, %0, %1, %2;" : "+f"(r[k]) : "f"(r[(k + 2) % CHAINS]), "f"(a)); }   // FFMA : FMNMX3
                if (MODE == 8) r[k] = fabsf(r[k]) - fabsf(a);                             // FADD with modifiers
                if (MODE == 9) { if ((k & 3) == 3) h[k] = __hmax2(h[k], __hmax2(h[(k + 4) % CHAINS], ha)); else h[k] = __hfma2(h[k], ha, hb); }   // 3 HFMA2 : 1 VHMNMX
                if (MODE == 10) h[k] = __hfma2(h[k], __floats2half2_rn(1.0f, 0.5f), hb);   // HFMA2 immediate form
                if (MODE == 11) h[k] = __hsub2(__habs2(h[k]), __habs2(ha));               // HADD2 |a| - |b|
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += r[k] + __low2float(h[k]) + __high2float(h[k]);
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, float* out, const float* cf, int sms, double mhz)
{
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(out, 64, cf);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, iters, cf);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)blocks * 8 * iters * 4 * CHAINS;       // warp instructions
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM  (%.3f per SMSP)\n", name, ms, winstr / clk / sms, winstr / clk / sms / 4);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (max)\n", p.name, p.multiProcessorCount, mhz);
    float* out; cudaMalloc(&out, 1 << 24);
    float h[4] = { 0.999f, 0.001f, 0.998f, 1.001f }; float* cf; cudaMalloc(&cf, 16); cudaMemcpy(cf, h, 16, cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA", out, cf, sms, mhz);
    run<1>("HFMA2", out, cf, sms, mhz);
    run<10>("HFMA2 imm", out, cf, sms, mhz);
    run<2>("HADD2", out, cf, sms, mhz);
    run<11>("HADD2 |a|-|b|", out, cf, sms, mhz);
    run<3>("VHMNMX 3-input", out, cf, sms, mhz);
    run<4>("FMNMX3", out, cf, sms, mhz);
    run<8>("FADD |a|-|b|", out, cf, sms, mhz);
    run<5>("FFMA:HFMA2 1:1", out, cf, sms, mhz);
    run<6>("HFMA2:VHMNMX 1:1", out, cf, sms, mhz);
    run<9>("HFMA2:VHMNMX 3:1", out, cf, sms, mhz);
    run<7>("FFMA:FMNMX3 1:1", out, cf, sms, mhz);
    return 0;
}
