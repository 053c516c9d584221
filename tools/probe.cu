// Micro-probes for FP64 latency / constant-operand throughput on sm_100a (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe tools/probe.cu && ./probe
#include <cstdio>
#include <cuda_runtime.h>
struct Tri { double v[528]; };

template <int ILP, bool FMA>
__global__ void k_chain(double* out, int iters, long long* cyc) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = FMA ? __fma_rn(a[i], b, c) : __dadd_rn(a[i], c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// emulates the sweep inner loop: acc = acc + M[e]*x  with M from the constant bank (kernel param), streaming through 4 KB
template <int ILP, bool EXACT>
__global__ void __launch_bounds__(1024) k_const(double* out, int reps, long long* cyc, const __grid_constant__ Tri M) {
    double a[ILP], x[ILP];
    // desynchronise the warps like real work would: warp w idles w*300 cycles first
    { long long s0 = clock64(); while (clock64() - s0 < (threadIdx.x >> 5) * 300) {} }
    for (int i = 0; i < ILP; ++i) { a[i] = 0.0; x[i] = threadIdx.x * 1e-3 + i; }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int e = 0; e < 528; ++e) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = EXACT ? __dadd_rn(a[i], __dmul_rn(M.v[e], x[i])) : __fma_rn(M.v[e], x[i], a[i]);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// same but the matrix comes from shared memory (broadcast LDS)
template <int ILP, bool EXACT>
__global__ void __launch_bounds__(1024) k_smem(double* out, int reps, long long* cyc, const double* Mg) {
    __shared__ double M[528];
    for (int i = threadIdx.x; i < 528; i += blockDim.x) M[i] = Mg[i];
    __syncthreads();
    double a[ILP], x[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = 0.0; x[i] = threadIdx.x * 1e-3 + i; }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int e = 0; e < 528; e += 2) {
            const double2 m2 = *reinterpret_cast<const double2*>(&M[e]);
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = EXACT ? __dadd_rn(a[i], __dmul_rn(m2.x, x[i])) : __fma_rn(m2.x, x[i], a[i]);
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = EXACT ? __dadd_rn(a[i], __dmul_rn(m2.y, x[i])) : __fma_rn(m2.y, x[i], a[i]);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename F> void run(const char* name, F launch, double ops_per_thread_iter, int warps_per_sm) {
    long long* cyc; double* out; cudaMalloc(&cyc, 8); cudaMalloc(&out, 8);
    launch(out, cyc); cudaError_t er = cudaDeviceSynchronize(); if (er != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(er)); return; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); launch(out, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SM %2d: %8.2f cycles per FP64 op-group (block 0), %.3f ms\n", name, warps_per_sm, double(h) / ops_per_thread_iter, ms);
}

int main() {
    Tri M; for (int i = 0; i < 528; ++i) M.v[i] = 1.0 / (1 + i);
    double* Mg; cudaMalloc(&Mg, sizeof(M)); cudaMemcpy(Mg, &M, sizeof(M), cudaMemcpyHostToDevice);
    const int it = 20000;
    for (int w : {1, 4, 8}) {
        int thr = 32 * w;  // one block per SM, w warps -> w/4 per SMSP (w=1: a lone warp)
        run("DFMA dependent chain ILP1", [&](double* o, long long* c) { k_chain<1, true><<<148, thr>>>(o, it, c); }, it, w);
        run("DADD dependent chain ILP1", [&](double* o, long long* c) { k_chain<1, false><<<148, thr>>>(o, it, c); }, it, w);
        run("DFMA chains ILP4 (per 4 ops)", [&](double* o, long long* c) { k_chain<4, true><<<148, thr>>>(o, it, c); }, it, w);
    }
    for (int w : {4, 8, 12, 24, 32}) {
        int thr = 32 * w > 1024 ? 1024 : 32 * w; int blocks = 148 * ((32 * w + thr - 1) / thr);
        run("const-bank  EXACT ILP1 (per term)", [&](double* o, long long* c) { k_const<1, true><<<blocks, thr>>>(o, 20, c, M); }, 20 * 528.0, w);
        run("const-bank  EXACT ILP2 (per 2 terms)", [&](double* o, long long* c) { k_const<2, true><<<blocks, thr>>>(o, 20, c, M); }, 20 * 528.0, w);
        run("const-bank  FMA   ILP2 (per 2 terms)", [&](double* o, long long* c) { k_const<2, false><<<blocks, thr>>>(o, 20, c, M); }, 20 * 528.0, w);
        run("smem-bcast  EXACT ILP1 (per term)", [&](double* o, long long* c) { k_smem<1, true><<<blocks, thr>>>(o, 20, c, Mg); }, 20 * 528.0, w);
        run("smem-bcast  EXACT ILP2 (per 2 terms)", [&](double* o, long long* c) { k_smem<2, true><<<blocks, thr>>>(o, 20, c, Mg); }, 20 * 528.0, w);
        run("smem-bcast  FMA   ILP2 (per 2 terms)", [&](double* o, long long* c) { k_smem<2, false><<<blocks, thr>>>(o, 20, c, Mg); }, 20 * 528.0, w);
        run("smem-bcast  FMA   ILP4 (per 4 terms)", [&](double* o, long long* c) { k_smem<4, false><<<blocks, thr>>>(o, 20, c, Mg); }, 20 * 528.0, w);
    }
    return 0;
}
