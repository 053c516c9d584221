// In-situ probe of the register-fibre triangle code: smem -> regs -> apply -> smem, no global traffic.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define LPFNT_MAXT 32
#include "../lpfun_b200/csrc/sweep_kernels.cuh"
using namespace lpfnt;

template <int MAXT, int PAD, bool FMA, int ITEMS>
__global__ void __launch_bounds__(128) k_probe(double* out, int reps, int tlen, int spread, const __grid_constant__ PackedTri<MAXT> M) {
    extern __shared__ double buf[];   // [ITEMS*128 rows? ] we use 128*ITEMS columns-fibres: fibre f at column (f%PAD'), ...
    const int tid = threadIdx.x;
    // each thread owns ITEMS columns of a [MAXT x (128*ITEMS)] array -> conflict-free stride access
    const int W = 128 * ITEMS;
    for (int i = tid; i < MAXT * W; i += 128) buf[i] = 1.0 / (1 + i);
    __syncthreads();
    // fibre length: uniform tlen, or spread across warps like a sorted tile (warp w gets tlen - w*spread)
    const int t = max(1, tlen - (tid >> 5) * spread);
    for (int r = 0; r < reps; ++r) {
        if constexpr (ITEMS == 1) {
            double x[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[k * W + tid] : 0.0;
            apply_fibre<MAXT, kLowerMatvec, FMA>(x, t, M);
#pragma unroll
            for (int k = 0; k < MAXT; ++k) if (k < t) buf[k * W + tid] = x[k];
        } else {
            double x[MAXT], y[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) { x[k] = (k < t) ? buf[k * W + tid] : 0.0; y[k] = (k < t) ? buf[k * W + 128 + tid] : 0.0; }
            apply_fibre2<MAXT, kLowerMatvec, FMA>(x, y, t, M);
#pragma unroll
            for (int k = 0; k < MAXT; ++k) if (k < t) { buf[k * W + tid] = x[k]; buf[k * W + 128 + tid] = y[k]; }
        }
        __syncthreads();
    }
    if (buf[tid] == 1.2345) out[0] = buf[tid];
}

template <bool FMA, int ITEMS>
void run(const char* name, int ctas_per_sm, int tlen, int spread) {
    PackedTri<32> M; for (int i = 0; i < 528; ++i) M.v[i] = 1e-3 / (1 + i);
    double* out; cudaMalloc(&out, 8);
    auto kern = k_probe<32, 31, FMA, ITEMS>;
    size_t smem = sizeof(double) * 32 * 128 * ITEMS;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 200;
    kern<<<148 * ctas_per_sm, 128, smem>>>(out, 2, tlen, spread, M); 
    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); kern<<<148 * ctas_per_sm, 128, smem>>>(out, reps, tlen, spread, M); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // useful terms: per thread per rep ITEMS * t(t+1)/2 with per-warp t
    double terms = 0; for (int w = 0; w < 4; ++w) { int t = tlen - w * spread; if (t < 1) t = 1; terms += 32.0 * ITEMS * t * (t + 1) / 2; }
    terms *= double(reps) * 148 * ctas_per_sm;
    double peak = 148.0 * 64 * 1.965e9 * (FMA ? 1.0 : 0.5);   // terms/s at the FP64 pipe peak (nominal clock)
    printf("%-12s items/thr %d  CTAs/SM %d  t=%2d spread=%d : %7.3f ms  %6.2f Tterm/s  = %5.1f %% of FP64 pipe peak\n", name, ITEMS, ctas_per_sm, tlen, spread, ms, terms / ms / 1e9, 100.0 * terms / (ms * 1e-3) / peak);
}

int main() {
    for (int c : {1, 2, 3, 6}) { run<false, 1>("exact", c, 31, 0); run<true, 1>("fma", c, 31, 0); }
    for (int c : {1, 2, 3}) { run<false, 2>("exact", c, 31, 0); run<true, 2>("fma", c, 31, 0); }
    run<false, 1>("exact", 6, 31, 7); run<true, 1>("fma", 6, 31, 7);
    run<false, 2>("exact", 3, 31, 7); run<true, 2>("fma", 3, 31, 7);
    run<false, 1>("exact", 6, 12, 0); run<false, 2>("exact", 3, 12, 0);
    return 0;
}
