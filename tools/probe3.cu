// Where do the matrix constants of the triangle code come from fastest?  Registers-only lower mat-vec
// (the 4-row block code of apply_fibre, no shared/global data traffic), constants through
//   V0 kernel parameter, static offsets (LDCU -> uniform-register operand: what the kernels use),
//   V1 __constant__ array with a per-thread (run-time zero) offset (LDC -> ordinary register operand),
//   V2 every second constant through V0 / V1,
//   V3 __constant__ array, static offsets.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int T = 32;
struct Tri { double v[T * (T + 1) / 2]; };
__constant__ double cM[T * (T + 1) / 2];

template <bool FMA> __device__ __forceinline__ double mac(double a, double l, double x) {
    if constexpr (FMA) return fma(l, x, a);
    else return __dadd_rn(a, __dmul_rn(l, x));
}

template <int VAR> __device__ __forceinline__ double cst(const Tri &M, int idx, int zz, int sel) {
    if (VAR == 0) return M.v[idx];
    if (VAR == 1) return cM[idx + zz];
    if (VAR == 2) return (sel & 1) ? M.v[idx] : cM[idx + zz];
    return cM[idx];
}

template <int VAR, bool FMA>
__global__ void __launch_bounds__(128, 6) k_probe(double *out, int reps, int z, const __grid_constant__ Tri M) {
    double x[T];
#pragma unroll
    for (int k = 0; k < T; ++k) x[k] = 1.0 / (1 + k + threadIdx.x);
    const int zz = (threadIdx.x + blockIdx.x) * z;  // 0 at run time, unknown to the compiler, per thread
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = T - 1; k >= 3; k -= 4) {
            double a3 = 0.0, a2 = 0.0, a1 = 0.0, a0 = 0.0;
#pragma unroll
            for (int j = 0; j <= k - 3; ++j) {
                a3 = mac<FMA>(a3, cst<VAR>(M, k * (k + 1) / 2 + j, zz, j), x[j]);
                a2 = mac<FMA>(a2, cst<VAR>(M, (k - 1) * k / 2 + j, zz, j + 1), x[j]);
                a1 = mac<FMA>(a1, cst<VAR>(M, (k - 2) * (k - 1) / 2 + j, zz, j), x[j]);
                a0 = mac<FMA>(a0, cst<VAR>(M, (k - 3) * (k - 2) / 2 + j, zz, j + 1), x[j]);
            }
            a3 = mac<FMA>(a3, cst<VAR>(M, k * (k + 1) / 2 + k - 2, zz, 0), x[k - 2]);
            a2 = mac<FMA>(a2, cst<VAR>(M, (k - 1) * k / 2 + k - 2, zz, 1), x[k - 2]);
            a1 = mac<FMA>(a1, cst<VAR>(M, (k - 2) * (k - 1) / 2 + k - 2, zz, 0), x[k - 2]);
            a3 = mac<FMA>(a3, cst<VAR>(M, k * (k + 1) / 2 + k - 1, zz, 1), x[k - 1]);
            a2 = mac<FMA>(a2, cst<VAR>(M, (k - 1) * k / 2 + k - 1, zz, 0), x[k - 1]);
            a3 = mac<FMA>(a3, cst<VAR>(M, k * (k + 1) / 2 + k, zz, 1), x[k]);
            x[k] = a3; x[k - 1] = a2; x[k - 2] = a1; x[k - 3] = a0;
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) s += x[k];
    if (s == 1.2345) out[0] = s;
}

template <int VAR, bool FMA>
void run(const char *name) {
    Tri M;
    for (int i = 0; i < T * (T + 1) / 2; ++i) M.v[i] = 1e-3 / (1 + i);
    cudaMemcpyToSymbol(cM, M.v, sizeof(M.v));
    double *out;
    cudaMalloc(&out, 8);
    const int reps = 400, grid = 148 * 6;
    k_probe<VAR, FMA><<<grid, 128>>>(out, 2, 0, M);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_probe<VAR, FMA><<<grid, 128>>>(out, reps, 0, M);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double terms = double(T) * (T + 1) / 2 * 128.0 * grid * reps;
    const double peak = 148.0 * 64 * 1.965e9 * (FMA ? 1.0 : 0.5);
    printf("%-36s %s : %7.3f ms  %6.2f Tterm/s = %5.1f %% of the FP64 pipe\n", name, FMA ? "fma  " : "exact", ms, terms / ms / 1e9, 100.0 * terms / (ms * 1e-3) / peak);
}

int main() {
    run<0, true>("V0 param, LDCU -> UR operand"); run<0, false>("V0 param, LDCU -> UR operand");
    run<1, true>("V1 __constant__ + reg offset (LDC)"); run<1, false>("V1 __constant__ + reg offset (LDC)");
    run<2, true>("V2 half/half"); run<2, false>("V2 half/half");
    run<3, true>("V3 __constant__ static"); run<3, false>("V3 __constant__ static");
    return 0;
}
