// Round-2 micro-benchmarks of the FP64 triangle code on B200 (registers / shared memory only, no HBM traffic):
//   F*  pipe peaks: DFMA, DMUL+DADD pairs, with register and with uniform-register (kernel parameter) operands
//   A   the shipped row form (apply_fibre, ROWS = 2 / 4), one fibre per thread, constants streamed by LDCU
//   B   two fibres per thread sharing the constants (apply_fibre2)
//   C   column-outer accumulate form: x streamed from shared memory, up to 32 accumulators per thread
//   D   FP64 tensor-core MMA (mma.sync.m8n8k4.f64): rate alone and mixed with DFMA; bit comparison with an FMA chain
//   E   constants from shared memory (LDS.128 broadcast) instead of the constant bank
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe4 tools/probe4.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int T = 32;
struct Tri { double v[T * (T + 1) / 2]; };

template <bool FMA> __device__ __forceinline__ double mac(double a, double l, double x) {
    if constexpr (FMA) return __fma_rn(l, x, a);
    else return __dadd_rn(a, __dmul_rn(l, x));
}

// ---------------------------------------------------------------- F: pipe peaks
template <int KIND>  // 0 DFMA reg, 1 DMUL+DADD reg, 2 DFMA with UR operand, 3 DMUL(UR)+DADD
__global__ void __launch_bounds__(256) k_peak(double *out, int iters, const __grid_constant__ Tri M) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x + i;
    const double b = 1.0000001 + 1e-9 * threadIdx.x, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) a[i] = __fma_rn(a[i], b, c);
            if (KIND == 1) a[i] = __dadd_rn(c, __dmul_rn(a[i], b));
            if (KIND == 2) a[i] = __fma_rn(a[i], M.v[i], c);
            if (KIND == 3) a[i] = __dadd_rn(c, __dmul_rn(a[i], M.v[i]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
}

// ---------------------------------------------------------------- A / B: row forms, registers only
template <bool FMA, int ROWS>
__device__ __forceinline__ void tri_rows(double (&x)[T], const Tri &M, const int tg) {
#pragma unroll
    for (int k = T - 1; k >= ROWS - 1; k -= ROWS) {
        if (k - (ROWS - 1) >= tg) continue;  // warp-uniform guard, as in the shipped kernels (also a scheduling barrier)
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

template <bool FMA, int ROWS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_rows(double *out, int reps, int tg, const __grid_constant__ Tri M) {
    double x[T];
#pragma unroll
    for (int k = 0; k < T; ++k) x[k] = 1.0 / (1 + k + threadIdx.x);
#pragma unroll 1
    for (int r = 0; r < reps; ++r) tri_rows<FMA, ROWS>(x, M, tg);
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k];
    if (s == 1.2345) out[0] = s;
}
template <bool FMA, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_rows2(double *out, int reps, int tg, const __grid_constant__ Tri M) {
    double x[T], y[T];
#pragma unroll
    for (int k = 0; k < T; ++k) { x[k] = 1.0 / (1 + k + threadIdx.x); y[k] = 2.0 / (3 + k + threadIdx.x); }
#pragma unroll 1
    for (int r = 0; r < reps; ++r) tri_rows2<FMA>(x, y, M, tg);
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k] + y[k];
    if (s == 1.2345) out[0] = s;
}

// ---------------------------------------------------------------- C: column-outer accumulate form, x streamed from smem
// F fibres per thread, all T rows accumulated (F*T accumulators), one pass over j; constants used F times each.
template <bool FMA, int F, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_cols(double *out, int reps, int tg, const __grid_constant__ Tri M) {
    extern __shared__ double buf[];  // [T][F*THREADS]
    const int W = F * THREADS;
    for (int i = threadIdx.x; i < T * W; i += THREADS) buf[i] = 1.0 / (1 + i);
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        double acc[F][T];
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
            for (int k = 0; k < T; ++k) acc[f][k] = 0.0;
#pragma unroll
        for (int j = 0; j < T; ++j) {
            if (j >= tg) continue;
            double xj[F];
#pragma unroll
            for (int f = 0; f < F; ++f) xj[f] = buf[j * W + f * THREADS + threadIdx.x];
#pragma unroll
            for (int k = j; k < T; ++k)
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f][k] = mac<FMA>(acc[f][k], M.v[k * (k + 1) / 2 + j], xj[f]);
        }
        __syncthreads();
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
            for (int k = 0; k < T; ++k) buf[k * W + f * THREADS + threadIdx.x] = acc[f][k] * 1e-3;
        __syncthreads();
    }
    if (buf[threadIdx.x] == 1.2345) out[0] = buf[threadIdx.x];
}

// ---------------------------------------------------------------- E: constants from shared memory (broadcast LDS.128)
template <bool FMA, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_rows_lds(double *out, int reps, int tg, const __grid_constant__ Tri M) {
    __shared__ __align__(16) double sM[T * (T + 1) / 2];
    for (int i = threadIdx.x; i < T * (T + 1) / 2; i += THREADS) sM[i] = M.v[i];
    __syncthreads();
    double x[T];
#pragma unroll
    for (int k = 0; k < T; ++k) x[k] = 1.0 / (1 + k + threadIdx.x);
    const volatile double *vm = sM;
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = T - 1; k >= 1; k -= 2) {
            if (k - 1 >= tg) continue;
            double a1 = 0.0, a0 = 0.0;
#pragma unroll
            for (int j = 0; j < k; ++j) {
                a1 = mac<FMA>(a1, sM[k * (k + 1) / 2 + j], x[j]);
                a0 = mac<FMA>(a0, sM[(k - 1) * k / 2 + j], x[j]);
            }
            x[k] = mac<FMA>(a1, sM[k * (k + 1) / 2 + k], x[k]);
            x[k - 1] = a0;
        }
        if (vm[0] == 1.2345) x[0] += 1.0;  // keeps the compiler from hoisting the LDS out of the rep loop
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k];
    if (s == 1.2345) out[0] = s;
}

// ---------------------------------------------------------------- D: FP64 MMA
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MIX>  // MIX = DFMA instructions issued per DMMA (0: DMMA only)
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters) {
    double d[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; }
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 1.0 + i + threadIdx.x;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dmma(d[i][0], d[i][1], a, b);
#pragma unroll
            for (int q = 0; q < MIX; ++q) f[(i + q) & 7] = __fma_rn(f[(i + q) & 7], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + f[i];
    if (s == 1.2345) out[0] = s;
}
// one 8x8x4 tile: DMMA against a sequential FMA chain (k ascending), bit compare
__global__ void k_dmma_bits(const double *A, const double *B, const double *C, double *Dm, double *Df) {
    const int lane = threadIdx.x;
    const int g = lane >> 2, tg = lane & 3;
    double a = A[g * 4 + tg];          // A row-major 8x4: a = A[row g][col tg]
    double b = B[tg * 8 + g];          // B 4x8 (k x n): b = B[k tg][n g]
    double c0 = C[g * 8 + 2 * tg], c1 = C[g * 8 + 2 * tg + 1];
    dmma(c0, c1, a, b);
    Dm[g * 8 + 2 * tg] = c0; Dm[g * 8 + 2 * tg + 1] = c1;
    if (lane < 32) {
        for (int e = lane; e < 64; e += 32) {
            const int i = e / 8, n = e % 8;
            double acc = C[e];
            for (int k = 0; k < 4; ++k) acc = __fma_rn(A[i * 4 + k], B[k * 8 + n], acc);
            Df[e] = acc;
        }
    }
}

// ---------------------------------------------------------------- host
static double g_peak_fma = 0.0;  // FMA/s measured by F0
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

void report(const char *name, double terms, bool fma, float ms) {
    const double rate = terms / (ms * 1e-3);
    const double peak = fma ? g_peak_fma : g_peak_fma * 0.5;
    printf("%-64s %s %8.3f ms %7.2f Tterm/s = %5.1f %% of the measured FP64 pipe\n", name, fma ? "fma  " : "exact", ms, rate / 1e12, 100.0 * rate / peak);
    fflush(stdout);
}

int main() {
    Tri M;
    for (int i = 0; i < T * (T + 1) / 2; ++i) M.v[i] = 1e-3 / (1 + i);
    double *out; cudaMalloc(&out, 8);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int SM = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, SM);
    {   // F: peaks
        const int it = 20000;
        const double ops = double(SM) * 8 * 256 * 8.0 * it;
        float ms = time_kernel(k_peak<0>, dim3(SM * 8), dim3(256), 0, out, it, M);
        g_peak_fma = ops / (ms * 1e-3);
        printf("F0 DFMA register operands: %.3f ms -> %.2f TFMA/s = %.1f FMA/clk/SM at 1965 MHz\n", ms, g_peak_fma / 1e12, g_peak_fma / SM / 1.965e9);
        ms = time_kernel(k_peak<1>, dim3(SM * 8), dim3(256), 0, out, it, M);
        printf("F1 DMUL+DADD register operands: %.3f ms -> %.2f Tpair/s (%.1f %% of DFMA/2)\n", ms, ops / (ms * 1e-3) / 1e12, 100.0 * ops / (ms * 1e-3) / (0.5 * g_peak_fma));
        ms = time_kernel(k_peak<2>, dim3(SM * 8), dim3(256), 0, out, it, M);
        printf("F2 DFMA uniform-register operand: %.3f ms -> %.2f TFMA/s\n", ms, ops / (ms * 1e-3) / 1e12);
        ms = time_kernel(k_peak<3>, dim3(SM * 8), dim3(256), 0, out, it, M);
        printf("F3 DMUL(UR)+DADD: %.3f ms -> %.2f Tpair/s\n", ms, ops / (ms * 1e-3) / 1e12);
    }
    const double tri = double(T) * (T + 1) / 2;
    const int reps = 300;
#define RUN_ROWS(FMA, ROWS, THREADS, MINB, CTAS)                                                                  \
    {                                                                                                               \
        float ms = time_kernel(k_rows<FMA, ROWS, THREADS, MINB>, dim3(SM * CTAS), dim3(THREADS), 0, out, reps, 32, M); \
        char nm[128]; snprintf(nm, 128, "A rows=%d, 1 fibre/thread, %d thr x %d CTA/SM", ROWS, THREADS, CTAS);     \
        report(nm, tri * THREADS * double(SM * CTAS) * reps, FMA, ms);                                              \
    }
    RUN_ROWS(false, 2, 768, 1, 1) RUN_ROWS(true, 2, 768, 1, 1)
    RUN_ROWS(false, 4, 768, 1, 1) RUN_ROWS(true, 4, 768, 1, 1)
    RUN_ROWS(false, 2, 384, 2, 2) RUN_ROWS(true, 2, 384, 2, 2)
    RUN_ROWS(false, 2, 384, 1, 1) RUN_ROWS(true, 2, 384, 1, 1)
    RUN_ROWS(false, 4, 384, 1, 1) RUN_ROWS(true, 4, 384, 1, 1)
    RUN_ROWS(false, 8, 384, 1, 1) RUN_ROWS(true, 8, 384, 1, 1)
    RUN_ROWS(false, 2, 128, 1, 1) RUN_ROWS(true, 2, 128, 1, 1)
    RUN_ROWS(false, 4, 128, 1, 1) RUN_ROWS(true, 4, 128, 1, 1)
#define RUN_ROWS2(FMA, THREADS, MINB, CTAS)                                                                    \
    {                                                                                                            \
        float ms = time_kernel(k_rows2<FMA, THREADS, MINB>, dim3(SM * CTAS), dim3(THREADS), 0, out, reps, 32, M);   \
        char nm[128]; snprintf(nm, 128, "B 2 fibres/thread (rows=2), %d thr x %d CTA/SM", THREADS, CTAS);       \
        report(nm, 2 * tri * THREADS * double(SM * CTAS) * reps, FMA, ms);                                       \
    }
    RUN_ROWS2(false, 384, 1, 1) RUN_ROWS2(true, 384, 1, 1)
    RUN_ROWS2(false, 192, 2, 2) RUN_ROWS2(true, 192, 2, 2)
    RUN_ROWS2(false, 128, 1, 1) RUN_ROWS2(true, 128, 1, 1)
#define RUN_COLS(FMA, F, THREADS, MINB, CTAS)                                                                          \
    {                                                                                                                    \
        auto kern = k_cols<FMA, F, THREADS, MINB>;                                                                       \
        const size_t smem = sizeof(double) * T * F * THREADS;                                                            \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));                              \
        float ms = time_kernel(kern, dim3(SM * CTAS), dim3(THREADS), smem, out, reps, 32, M);                                \
        char nm[128]; snprintf(nm, 128, "C accumulate form, %d fibre(s)/thread x 32 rows, %d thr x %d CTA/SM", F, THREADS, CTAS); \
        report(nm, F * tri * THREADS * double(SM * CTAS) * reps, FMA, ms);                                               \
    }
    RUN_COLS(false, 1, 384, 1, 1) RUN_COLS(true, 1, 384, 1, 1)
    RUN_COLS(false, 1, 256, 2, 2) RUN_COLS(true, 1, 256, 2, 2)
    RUN_COLS(false, 2, 192, 1, 1) RUN_COLS(true, 2, 192, 1, 1)
    RUN_COLS(false, 2, 128, 2, 2) RUN_COLS(true, 2, 128, 2, 2)
#define RUN_LDS(FMA, THREADS, MINB, CTAS)                                                                         \
    {                                                                                                               \
        float ms = time_kernel(k_rows_lds<FMA, THREADS, MINB>, dim3(SM * CTAS), dim3(THREADS), 0, out, reps, 32, M);   \
        char nm[128]; snprintf(nm, 128, "E constants by LDS broadcast (rows=2), %d thr x %d CTA/SM", THREADS, CTAS); \
        report(nm, tri * THREADS * double(SM * CTAS) * reps, FMA, ms);                                              \
    }
    RUN_LDS(false, 768, 1, 1) RUN_LDS(true, 768, 1, 1)
    {   // D: DMMA
        const int it = 20000;
        const double mmas = double(SM) * 8 * 8 * 8.0 * it;  // warps * 8 per iteration
        float ms = time_kernel(k_dmma<0>, dim3(SM * 8), dim3(256), 0, out, it);
        const double fmas = mmas * 256.0;
        printf("D0 mma.sync.m8n8k4.f64 alone: %.3f ms -> %.2f TFMA/s = %.1f %% of the DFMA rate\n", ms, fmas / (ms * 1e-3) / 1e12, 100.0 * fmas / (ms * 1e-3) / g_peak_fma);
        ms = time_kernel(k_dmma<4>, dim3(SM * 8), dim3(256), 0, out, it);
        const double tot = mmas * 256.0 + mmas * 4 * 32.0;
        printf("D1 DMMA + 4 DFMA each: %.3f ms -> DMMA part %.2f TFMA/s, DFMA part %.2f TFMA/s, sum %.1f %% of the DFMA rate\n", ms,
               mmas * 256.0 / (ms * 1e-3) / 1e12, mmas * 128.0 / (ms * 1e-3) / 1e12, 100.0 * tot / (ms * 1e-3) / g_peak_fma);
        ms = time_kernel(k_dmma<8>, dim3(SM * 8), dim3(256), 0, out, it);
        const double tot8 = mmas * 256.0 + mmas * 8 * 32.0;
        printf("D2 DMMA + 8 DFMA each: %.3f ms -> sum %.1f %% of the DFMA rate\n", ms, 100.0 * tot8 / (ms * 1e-3) / g_peak_fma);
        // bits
        std::vector<double> hA(32), hB(32), hC(64), hDm(64), hDf(64);
        srand(7);
        for (auto &v : hA) v = (rand() / double(RAND_MAX) - 0.5) * 1e3;
        for (auto &v : hB) v = (rand() / double(RAND_MAX) - 0.5);
        for (auto &v : hC) v = (rand() / double(RAND_MAX) - 0.5) * 1e-3;
        double *dA, *dB, *dC, *dDm, *dDf;
        cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512); cudaMalloc(&dDm, 512); cudaMalloc(&dDf, 512);
        cudaMemcpy(dA, hA.data(), 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), 256, cudaMemcpyHostToDevice); cudaMemcpy(dC, hC.data(), 512, cudaMemcpyHostToDevice);
        k_dmma_bits<<<1, 32>>>(dA, dB, dC, dDm, dDf);
        cudaMemcpy(hDm.data(), dDm, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hDf.data(), dDf, 512, cudaMemcpyDeviceToHost);
        int same = 0; double maxd = 0;
        for (int i = 0; i < 64; ++i) { same += memcmp(&hDm[i], &hDf[i], 8) == 0; maxd = fmax(maxd, fabs(hDm[i] - hDf[i])); }
        printf("D3 DMMA vs sequential FMA chain (k ascending): %d / 64 bit-identical, max |diff| %.3e\n", same, maxd);
    }
    return 0;
}
