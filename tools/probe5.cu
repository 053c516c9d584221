// Round-2 micro-benchmarks, part 2 (B200): how fast can ONE warp per SM sub-partition issue FP64 work, and what does the
// length imbalance of the real item lists cost inside one barrier interval?
//   G  k_chain<KIND, CH>: CH independent accumulator chains per thread, W warps per sub-partition (128*W threads, 1 CTA/SM)
//      KIND 0 DFMA, 1 DMUL+DADD (products independent, adds chained), 2 DADD only
//   H  triangle code (apply_fibre form, ROWS = 2 / 4) with a per-warp row bound tg[w] taken from the strict length order of
//      d=3, n=30 (23 warps), one __syncthreads per repetition (the sweep barrier) -- against the same work spread evenly
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe5 tools/probe5.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int T = 32;
struct Tri { double v[T * (T + 1) / 2]; };

template <bool FMA> __device__ __forceinline__ double mac(double a, double l, double x) {
    if constexpr (FMA) return __fma_rn(l, x, a);
    else return __dadd_rn(a, __dmul_rn(l, x));
}

template <int KIND, int CH>
__global__ void __launch_bounds__(1024) k_chain(double *out, int iters, long long *cyc) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = 1.0 + threadIdx.x + i;
    const double b = 1.0000001 + 1e-9 * threadIdx.x, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (KIND == 0) a[i] = __fma_rn(a[i], b, c);
                if (KIND == 1) a[i] = __dadd_rn(a[i], __dmul_rn(b, c + u + i));
                if (KIND == 2) a[i] = __dadd_rn(a[i], c);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <bool FMA, int ROWS>
__device__ __forceinline__ void tri_rows(double (&x)[T], const Tri &M, const int tg) {
#pragma unroll
    for (int k = T - 1; k >= ROWS - 1; k -= ROWS) {
        if (k - (ROWS - 1) >= tg) continue;
        double acc[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = 0.0;
#pragma unroll
        for (int j = 0; j <= k; ++j) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                if (j <= k - r) acc[r] = mac<FMA>(acc[r], M.v[(k - r) * (k - r + 1) / 2 + j], x[j]);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) x[k - r] = acc[r];
    }
}

struct Bounds { int tg[24]; };

template <bool FMA, int ROWS, bool BAR>
__global__ void __launch_bounds__(768, 1) k_imb(double *out, int reps, const __grid_constant__ Bounds B, const __grid_constant__ Tri M, long long *cyc) {
    double x[T];
#pragma unroll
    for (int k = 0; k < T; ++k) x[k] = 1.0 / (1 + k + threadIdx.x);
    const int tg = B.tg[threadIdx.x >> 5];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        // the row bound behind an offset the compiler cannot prove to be zero (keeps the constant loads inside the loop)
        tri_rows<FMA, ROWS>(x, *reinterpret_cast<const Tri *>(&M.v[(r >> 20) << 1]), tg);
        if (BAR) __syncthreads();
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}


template <bool FMA>
__device__ __forceinline__ void tri_rows2(double (&x)[T], double (&y)[T], const Tri &M, const int tg) {
#pragma unroll
    for (int k = T - 1; k >= 1; k -= 2) {
        if (k - 1 >= tg) continue;
        double a1 = 0.0, a0 = 0.0, b1 = 0.0, b0 = 0.0;
#pragma unroll
        for (int j = 0; j < k; ++j) {
            const double m1 = M.v[k * (k + 1) / 2 + j], m0 = M.v[(k - 1) * k / 2 + j];
            a1 = mac<FMA>(a1, m1, x[j]); b1 = mac<FMA>(b1, m1, y[j]);
            a0 = mac<FMA>(a0, m0, x[j]); b0 = mac<FMA>(b0, m0, y[j]);
        }
        const double md = M.v[k * (k + 1) / 2 + k];
        x[k] = mac<FMA>(a1, md, x[k]); y[k] = mac<FMA>(b1, md, y[k]);
        x[k - 1] = a0; y[k - 1] = b0;
    }
}
struct Bounds12 { int tg[12]; };
template <bool FMA, bool BAR>
__global__ void __launch_bounds__(384, 1) k_imb2(double *out, int reps, const __grid_constant__ Bounds12 B, const __grid_constant__ Tri M, long long *cyc) {
    double x[T], y[T];
#pragma unroll
    for (int k = 0; k < T; ++k) { x[k] = 1.0 / (1 + k + threadIdx.x); y[k] = 2.0 / (3 + k + threadIdx.x); }
    const int tg = B.tg[threadIdx.x >> 5];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        tri_rows2<FMA>(x, y, *reinterpret_cast<const Tri *>(&M.v[(r >> 20) << 1]), tg);
        if (BAR) __syncthreads();
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k] + y[k];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}


// exact form, two rows, the additions as volatile asm: ptxas keeps their order, i.e. the two chains stay interleaved
// (the plain form compiles to ONE serial chain per warp at 80 registers: DMUL -> DADD -> DMUL -> ...)
__device__ __forceinline__ void vadd(double &acc, double p) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc) : "d"(p)); }
__device__ __forceinline__ void tri_rows_v(double (&x)[T], const Tri &M, const int tg) {
#pragma unroll
    for (int k = T - 1; k >= 1; k -= 2) {
        if (k - 1 >= tg) continue;
        double a1 = 0.0, a0 = 0.0;
#pragma unroll
        for (int j = 0; j < k; ++j) {
            const double p1 = __dmul_rn(M.v[k * (k + 1) / 2 + j], x[j]);
            const double p0 = __dmul_rn(M.v[(k - 1) * k / 2 + j], x[j]);
            vadd(a1, p1);
            vadd(a0, p0);
        }
        const double pd = __dmul_rn(M.v[k * (k + 1) / 2 + k], x[k]);
        vadd(a1, pd);
        x[k] = a1;
        x[k - 1] = a0;
    }
}
template <bool BAR>
__global__ void __launch_bounds__(768, 1) k_imbv(double *out, int reps, const __grid_constant__ Bounds B, const __grid_constant__ Tri M, long long *cyc) {
    double x[T];
#pragma unroll
    for (int k = 0; k < T; ++k) x[k] = 1.0 / (1 + k + threadIdx.x);
    const int tg = B.tg[threadIdx.x >> 5];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        tri_rows_v(x, *reinterpret_cast<const Tri *>(&M.v[(r >> 20) << 1]), tg);
        if (BAR) __syncthreads();
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename K, typename... Args>
float time_kernel(K kern, dim3 grid, dim3 block, size_t smem, Args... args) {
    kern<<<grid, block, smem>>>(args...);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return -1.f; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<grid, block, smem>>>(args...);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    Tri M;
    for (int i = 0; i < T * (T + 1) / 2; ++i) M.v[i] = 1e-3 / (1 + i);
    double *out; cudaMalloc(&out, 8);
    long long *cyc; cudaMalloc(&cyc, 8);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int SM = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, SM);
    const int it = 4000;
    auto show = [&](const char *kind, int ch, int w, long long c, double fp64_per_iter) {
        // pipe cycles: every FP64 warp instruction occupies its sub-partition's pipe for 2 cycles
        const double instr = fp64_per_iter * it * 8 * ch * w;  // per sub-partition
        printf("G %-10s chains=%d warps/subpartition=%d: %lld cycles, %.2f cycles per FP64 instruction of a warp, pipe busy %.1f %%\n", kind, ch, w, c,
               double(c) / (fp64_per_iter * it * 8 * ch), 100.0 * instr * 2 / double(c));
    };
#define RUN_G(KIND, NAME, CH, W, FPI)                                                      \
    {                                                                                       \
        k_chain<KIND, CH><<<SM, 128 * W>>>(out, it, cyc);                                   \
        cudaDeviceSynchronize();                                                            \
        k_chain<KIND, CH><<<SM, 128 * W>>>(out, it, cyc);                                   \
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                        \
        show(NAME, CH, W, c, FPI);                                                          \
    }
    RUN_G(0, "DFMA", 1, 1, 1) RUN_G(0, "DFMA", 2, 1, 1) RUN_G(0, "DFMA", 4, 1, 1) RUN_G(0, "DFMA", 8, 1, 1)
    RUN_G(0, "DFMA", 8, 2, 1) RUN_G(0, "DFMA", 8, 4, 1) RUN_G(0, "DFMA", 2, 6, 1)
    RUN_G(2, "DADD", 1, 1, 1) RUN_G(2, "DADD", 2, 1, 1) RUN_G(2, "DADD", 4, 1, 1) RUN_G(2, "DADD", 8, 1, 1)
    RUN_G(1, "DMUL+DADD", 1, 1, 2) RUN_G(1, "DMUL+DADD", 2, 1, 2) RUN_G(1, "DMUL+DADD", 4, 1, 2) RUN_G(1, "DMUL+DADD", 8, 1, 2)
    RUN_G(1, "DMUL+DADD", 2, 2, 2) RUN_G(1, "DMUL+DADD", 2, 3, 2) RUN_G(1, "DMUL+DADD", 2, 4, 2) RUN_G(1, "DMUL+DADD", 2, 6, 2)

    // H: strict length order of d=3, n=30: longest item of each chunk of 32 (23 chunks), dealt in snake order
    const int chunk_max[23] = {31, 30, 29, 29, 28, 27, 26, 26, 25, 24, 23, 22, 21, 21, 19, 18, 17, 16, 15, 13, 11, 9, 7};
    Bounds Bs{}, Bu{}, Bf{};
    long long work = 0;
    for (int c = 0; c < 24; ++c) {
        const int grp = c / 4, pos = c % 4;
        const int w = grp * 4 + ((grp & 1) ? 3 - pos : pos);
        Bs.tg[w] = c < 23 ? chunk_max[c] : 0;
        Bf.tg[c] = 32;
    }
    for (int w = 0; w < 24; ++w) { const int tt = (Bs.tg[w] + 1) & ~1; work += tt * (tt + 1) / 2; }
    // the same work spread evenly: every warp gets the bound whose triangle is work / 24
    int te = 2; while ((te + 2) * (te + 3) / 2 <= work / 24) te += 2;
    for (int w = 0; w < 24; ++w) Bu.tg[w] = te;
    printf("H strict-order bounds: executed terms per lane-slot %lld (sum over 24 warps), even bound %d (%d terms x 24 = %d)\n", work, te, te * (te + 1) / 2, 24 * te * (te + 1) / 2);
    const int reps = 200;
#define RUN_H(FMA, ROWS, BAR, BND, NAME)                                                                       \
    {                                                                                                           \
        k_imb<FMA, ROWS, BAR><<<SM, 768>>>(out, reps, BND, M, cyc);                                             \
        cudaDeviceSynchronize();                                                                                \
        k_imb<FMA, ROWS, BAR><<<SM, 768>>>(out, reps, BND, M, cyc);                                             \
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                            \
        long long wk = 0;                                                                                       \
        for (int w = 0; w < 24; ++w) { const int tt = ((BND.tg[w] + ROWS - 1) / ROWS) * ROWS; wk += tt * (tt + 1) / 2; } \
        const double pipe = double(wk) * (FMA ? 2.0 : 4.0) / 4.0;  /* pipe cycles per sub-partition and repetition */ \
        printf("H %-28s %s rows=%d barrier=%d: %8.0f cycles per sweep, pipe-time %6.0f -> %.1f %% busy\n", NAME, FMA ? "fma  " : "exact", ROWS, int(BAR), double(c) / reps, pipe, 100.0 * pipe * reps / double(c)); \
    }
    RUN_H(false, 2, true, Bf, "all warps 32 rows")
    RUN_H(false, 2, true, Bs, "strict order, snake")
    RUN_H(false, 2, false, Bs, "strict order, snake")
    RUN_H(false, 2, true, Bu, "same work, even")
    RUN_H(false, 4, true, Bs, "strict order, snake")
    RUN_H(false, 4, true, Bu, "same work, even")
    RUN_H(true, 2, true, Bf, "all warps 32 rows")
    RUN_H(true, 2, true, Bs, "strict order, snake")
    RUN_H(true, 2, true, Bu, "same work, even")
    RUN_H(true, 4, true, Bs, "strict order, snake")
    RUN_H(true, 4, true, Bu, "same work, even")


#define RUN_HV(BND, NAME)                                                                                      \
    {                                                                                                           \
        k_imbv<true><<<SM, 768>>>(out, reps, BND, M, cyc);                                                      \
        cudaDeviceSynchronize();                                                                                \
        k_imbv<true><<<SM, 768>>>(out, reps, BND, M, cyc);                                                      \
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                            \
        long long wk = 0;                                                                                       \
        for (int w = 0; w < 24; ++w) { const int tt = (BND.tg[w] + 1) & ~1; wk += tt * (tt + 1) / 2; }          \
        const double pipe = double(wk) * 4.0 / 4.0;                                                             \
        printf("HV %-27s exact rows=2, ordered additions: %8.0f cycles per sweep, pipe-time %6.0f -> %.1f %% busy\n", NAME, double(c) / reps, pipe, 100.0 * pipe * reps / double(c)); \
    }
    RUN_HV(Bf, "all warps 32 rows")
    RUN_HV(Bs, "strict order, snake")
    RUN_HV(Bu, "same work, even")
    {   // H2: two fibres per thread, 384 threads, 12 warps
        const int cm[12] = {31, 29, 28, 26, 25, 23, 21, 19, 17, 15, 11, 7};
        Bounds12 Cs{}, Cf{}, Cu{};
        long long wk = 0;
        for (int c = 0; c < 12; ++c) {
            const int grp = c / 4, pos = c % 4;
            const int w = grp * 4 + ((grp & 1) ? 3 - pos : pos);
            Cs.tg[w] = cm[c]; Cf.tg[c] = 32;
            const int tt = (cm[c] + 1) & ~1; wk += 2 * tt * (tt + 1) / 2;
        }
        int te = 2; while (2 * (te + 2) * (te + 3) / 2 * 12 <= wk) te += 2;
        for (int c = 0; c < 12; ++c) Cu.tg[c] = te;
#define RUN_H2(FMA, BAR, BND, NAME)                                                                            \
        {                                                                                                       \
            k_imb2<FMA, BAR><<<SM, 384>>>(out, reps, BND, M, cyc);                                              \
            cudaDeviceSynchronize();                                                                            \
            k_imb2<FMA, BAR><<<SM, 384>>>(out, reps, BND, M, cyc);                                              \
            cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) printf("failed %s\n", cudaGetErrorString(e)); \
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                        \
            long long w2 = 0;                                                                                   \
            for (int w = 0; w < 12; ++w) { const int tt = (BND.tg[w] + 1) & ~1; w2 += 2 * tt * (tt + 1) / 2; } \
            const double pipe = double(w2) * (FMA ? 2.0 : 4.0) / 4.0;                                           \
            printf("H2 %-27s %s 2 fibres/thread barrier=%d: %8.0f cycles per sweep, pipe-time %6.0f -> %.1f %% busy\n", NAME, FMA ? "fma  " : "exact", int(BAR), double(c) / reps, pipe, 100.0 * pipe * reps / double(c)); \
        }
        RUN_H2(false, true, Cf, "all warps 32 rows")
        RUN_H2(false, true, Cs, "strict order, snake")
        RUN_H2(false, true, Cu, "same work, even")
        RUN_H2(true, true, Cf, "all warps 32 rows")
        RUN_H2(true, true, Cs, "strict order, snake")
        RUN_H2(true, true, Cu, "same work, even")
    }
    return 0;
}
