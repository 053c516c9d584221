// Triangular sweeps along one coordinate of a lower-set vector (sm_100a).
//
// Work decomposition: ONE THREAD OWNS ONE ELEMENT-FIBRE -- the <= n+1 entries
// x[a | a_h -> 0..t-1] that couple along coordinate h.  The fibre lives in
// registers, the packed (n+1)x(n+1) triangle arrives as a __grid_constant__ kernel
// parameter, so every matrix entry is a constant-bank operand of the DMUL/DFMA that
// uses it (fully unrolled, no loads, no shared-memory traffic for the matrix).
//
// Arithmetic order is the reference's (src/lpfun/core/atoms.py
// itransform_lt_* :72,206,406,668,1022, dtransform_ut_md :1223, transform_lt_*
// :35,311,494,824): one left-to-right fold per output, ascending in the summation
// index, starting from +0.0.  EXACT rounds multiply and add separately
// (__dmul_rn/__dadd_rn are never contracted), FMA fuses them.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lpfnt {

enum SweepMode : int {
    kLowerMatvec = 0,  // out_k = sum_{j<=k} L[k,j] x_j                (fnt with inv V, ifnt, dxT)
    kUpperMatvec = 1,  // out_k = sum_{j>=k} U[k,j] x_j                (dx)
    kLowerSolve = 2,   // out_k = (x_k - sum_{j<k} L[k,j] out_j)/L[k,k] (fnt, precomputation=False)
    kUpperSolve = 3    // out_k = (x_k - (U[k,k]*0 + sum_{j>k} U[k,j] out_j))/U[k,k]
};

// Packed triangle with COMPILE-TIME offsets, independent of n:
//   lower kinds: v[k(k+1)/2 + j] = L[k][j]  (j <= k)
//   upper kinds: v[j(j+1)/2 + k] = U[k][j]  (k <= j)   (host re-packs the reference layout)
template <int MAXT>
struct PackedTri {
    double v[MAXT * (MAXT + 1) / 2];
};

template <bool FMA>
__device__ __forceinline__ double mac(double acc, double a, double b) {
    if constexpr (FMA) return __fma_rn(a, b, acc);
    else return __dadd_rn(acc, __dmul_rn(a, b));
}

// Apply the triangular operator to a register-resident fibre of length t, in place.
template <int MAXT, int MODE, bool FMA>
__device__ __forceinline__ void apply_fibre(double (&x)[MAXT], const int t, const PackedTri<MAXT> &M) {
    if constexpr (MODE == kLowerMatvec) {
        // Outputs from the top down: out_k only reads x_j, j <= k, so it may overwrite x_k.
        // Two rows are folded side by side (independent accumulators) to halve the dependent
        // DADD/DFMA chain a lone warp sees; each row keeps its own ascending-j order.
        static_assert(MAXT % 2 == 0, "rows are processed in pairs");
#pragma unroll
        for (int k = MAXT - 1; k >= 1; k -= 2) {
            if (k < t) {
                double a1 = 0.0, a0 = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j) {
                    a1 = mac<FMA>(a1, M.v[k * (k + 1) / 2 + j], x[j]);
                    a0 = mac<FMA>(a0, M.v[(k - 1) * k / 2 + j], x[j]);
                }
                a1 = mac<FMA>(a1, M.v[k * (k + 1) / 2 + k], x[k]);
                x[k] = a1;
                x[k - 1] = a0;
            } else if (k - 1 < t) {
                double a0 = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j) a0 = mac<FMA>(a0, M.v[(k - 1) * k / 2 + j], x[j]);
                x[k - 1] = a0;
            }
        }
    } else if constexpr (MODE == kUpperMatvec) {
        // column-outer: step j feeds x_j into out_0..out_j; each out_k still receives its
        // terms in ascending j.  out_j takes over the register of x_j at step j.
#pragma unroll
        for (int j = 0; j < MAXT; ++j) {
            if (j < t) {
                const double xj = x[j];
#pragma unroll
                for (int k = 0; k < j; ++k) x[k] = mac<FMA>(x[k], M.v[j * (j + 1) / 2 + k], xj);
                x[j] = mac<FMA>(0.0, M.v[j * (j + 1) / 2 + j], xj);
            }
        }
    } else if constexpr (MODE == kLowerSolve) {
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            if (k < t) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j) s = mac<false>(s, M.v[k * (k + 1) / 2 + j], x[j]);
                x[k] = __ddiv_rn(__dsub_rn(x[k], s), M.v[k * (k + 1) / 2 + k]);
            }
        }
    } else {  // kUpperSolve: backward substitution, rows from the last one up
#pragma unroll
        for (int k = MAXT - 1; k >= 0; --k) {
            if (k < t) {
                double s = mac<false>(0.0, M.v[k * (k + 1) / 2 + k], 0.0);
#pragma unroll
                for (int j = k + 1; j < MAXT; ++j)
                    if (j < t) s = mac<false>(s, M.v[j * (j + 1) / 2 + k], x[j]);
                x[k] = __ddiv_rn(__dsub_rn(x[k], s), M.v[k * (k + 1) / 2 + k]);
            }
        }
    }
}

struct LineSweepArgs {
    const double *in;
    double *out;
    const int32_t *line_off;  // N_1 + 1
    const int32_t *perm_in;   // gather: x[e] = in[perm_in[e]]  (or nullptr)
    const int32_t *perm_out;  // scatter: out[perm_out[e]] = y[e] (or nullptr)
    int64_t stride;           // elements between consecutive functions of the batch (= N)
    int32_t num_lines;
};

// Coordinate 0: the fibres are the contiguous lines.  A CTA stages the elements of
// blockDim.x consecutive lines in shared memory with coalesced (or permuted) loads, each
// thread pulls its own line into registers, sweeps it and puts it back; the store is
// coalesced again.  Grid: (ceil(N_1 / blockDim.x), batch).
template <int MAXT, int MODE, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS) k_sweep_lines(const LineSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    __shared__ double buf[THREADS * MAXT];
    const int l0 = blockIdx.x * THREADS;
    const int l1 = min(l0 + THREADS, a.num_lines);
    const int e0 = a.line_off[l0];
    const int cnt = a.line_off[l1] - e0;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;

    if (a.perm_in) {
        for (int i = threadIdx.x; i < cnt; i += THREADS) buf[i] = in[a.perm_in[e0 + i]];
    } else {
        for (int i = threadIdx.x; i < cnt; i += THREADS) buf[i] = in[e0 + i];
    }
    __syncthreads();
    const int l = l0 + threadIdx.x;
    if (l < l1) {
        const int o = a.line_off[l] - e0;
        const int t = a.line_off[l + 1] - a.line_off[l];
        double x[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[o + k] : 0.0;
        apply_fibre<MAXT, MODE, FMA>(x, t, M);
#pragma unroll
        for (int k = 0; k < MAXT; ++k)
            if (k < t) buf[o + k] = x[k];
    }
    __syncthreads();
    if (a.perm_out) {
        for (int i = threadIdx.x; i < cnt; i += THREADS) out[a.perm_out[e0 + i]] = buf[i];
    } else {
        for (int i = threadIdx.x; i < cnt; i += THREADS) out[e0 + i] = buf[i];
    }
}

struct FibreSweepArgs {
    const double *in;
    double *out;
    const int32_t *thr_fib;    // thread -> fibre
    const int32_t *thr_start;  // fibre -> first thread
    const int32_t *fib_ptr;    // fibre -> first entry
    const int2 *ent;           // {line element offset, line length}
    const int32_t *perm_out;
    int64_t stride;
    int32_t num_threads;
};

// Coordinate h >= 1: thread j owns element-fibre (fibre f, lane a_0); its entries sit at
// ent[k].off + a_0 for the first t lines with len > a_0 (line lengths never increase along
// a fibre).  Neighbouring lanes touch neighbouring addresses, so every load/store of a warp
// is a handful of contiguous runs.  Grid: (ceil(num_threads / blockDim.x), batch).
template <int MAXT, int MODE, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS) k_sweep_fibres(const FibreSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    const int j = blockIdx.x * THREADS + threadIdx.x;
    if (j >= a.num_threads) return;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    const int f = a.thr_fib[j];
    const int lane = j - a.thr_start[f];
    const int p0 = a.fib_ptr[f];
    const int nl = a.fib_ptr[f + 1] - p0;
    double x[MAXT];
    int t = 0;
#pragma unroll
    for (int k = 0; k < MAXT; ++k) {
        x[k] = 0.0;
        if (k < nl) {
            const int2 en = __ldg(a.ent + p0 + k);
            if (en.y > lane) {
                x[k] = in[en.x + lane];
                t = k + 1;
            }
        }
    }
    apply_fibre<MAXT, MODE, FMA>(x, t, M);
#pragma unroll
    for (int k = 0; k < MAXT; ++k) {
        if (k < t) {
            const int e = __ldg(a.ent + p0 + k).x + lane;
            out[a.perm_out ? a.perm_out[e] : e] = x[k];
        }
    }
}

struct TileSweepArgs {
    const double *in;
    double *out;
    const int32_t *tile_fib;   // tile -> first fibre
    const int32_t *fib_ptr;    // fibre -> first row (entry of ent)
    const int32_t *thr_start;  // fibre -> first item
    const int2 *ent;           // rows: {line element offset, line length}
    const uint32_t *items;     // element-fibres, per tile sorted by length: row | lane << 16 | t << 24
    const uint32_t *rows0;     // rows, per tile sorted by length: row | len << 24
    const int32_t *perm_in;    // colex gather on load (or nullptr)
    const int32_t *perm_out;   // colex scatter on store (or nullptr)
    int64_t stride;
};

// Row-tile sweep along coordinate h >= 1, optionally fused with the coordinate-0 sweep.
//
// A tile is a run of consecutive fibres of coordinate h: every fibre is a (a_0, a_h) slice made of
// <= n+1 contiguous lines.  The CTA copies those lines into shared memory (one warp per line:
// coalesced <= 248 B runs), row r at r*PAD with PAD odd, then
//   DIM0: thread <-> row, rows ordered by length   (coordinate 0: stride 1)
//   always: thread <-> element-fibre (slice, lane a_0), ordered by length (stride PAD)
// and writes the lines back.  Because the work items are ordered by length on the host, the 32
// lanes of a warp run triangles of nearly equal size (the unsorted thread<->lane mapping wastes
// 50-80 % of the FP64 lanes on l^2 sets).  Grid: (tiles, batch).
template <int MAXT, int MODE, bool FMA, bool DIM0, int THREADS>
__global__ void __launch_bounds__(THREADS) k_sweep_tiles(const TileSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    constexpr int PAD = MAXT + 1;
    __shared__ double buf[THREADS * PAD];
    const int f0 = a.tile_fib[blockIdx.x], f1 = a.tile_fib[blockIdx.x + 1];
    const int r0 = a.fib_ptr[f0], nrows = a.fib_ptr[f1] - r0;
    const int i0 = a.thr_start[f0], nitems = a.thr_start[f1] - i0;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int r = warp; r < nrows; r += THREADS / 32) {
        const int2 en = __ldg(a.ent + r0 + r);
        if (lane < en.y) {
            const int e = en.x + lane;
            buf[r * PAD + lane] = in[a.perm_in ? a.perm_in[e] : e];
        }
    }
    __syncthreads();
    if constexpr (DIM0) {
        if (threadIdx.x < nrows) {
            const uint32_t d = __ldg(a.rows0 + r0 + threadIdx.x);
            const int t = int(d >> 24);
            double *row = buf + int(d & 0xffffu) * PAD;
            double x[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? row[k] : 0.0;
            apply_fibre<MAXT, MODE, FMA>(x, t, M);
#pragma unroll
            for (int k = 0; k < MAXT; ++k)
                if (k < t) row[k] = x[k];
        }
        __syncthreads();
    }
    if (threadIdx.x < nitems) {
        const uint32_t d = __ldg(a.items + i0 + threadIdx.x);
        const int t = int(d >> 24);
        double *col = buf + int(d & 0xffffu) * PAD + int((d >> 16) & 0xffu);
        double x[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? col[k * PAD] : 0.0;
        apply_fibre<MAXT, MODE, FMA>(x, t, M);
#pragma unroll
        for (int k = 0; k < MAXT; ++k)
            if (k < t) col[k * PAD] = x[k];
    }
    __syncthreads();
    for (int r = warp; r < nrows; r += THREADS / 32) {
        const int2 en = __ldg(a.ent + r0 + r);
        if (lane < en.y) {
            const int e = en.x + lane;
            out[a.perm_out ? a.perm_out[e] : e] = buf[r * PAD + lane];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Generic fallback for fibres longer than the register kernels support (n + 1 > 32): same
// thread <-> element-fibre mapping, matrix read from global memory in the REFERENCE packing,
// fibre re-read from global memory (O(t^2) loads, L1/L2 resident).  `in` may equal `out`
// only when no permutation is applied.
// ---------------------------------------------------------------------------------------
struct GenericArgs {
    const double *in;
    double *out;
    const double *mat;         // reference packing: lower row k at k(k+1)/2; upper row k at k*n1 - k(k-1)/2
    const int32_t *line_off;
    const int32_t *thr_fib, *thr_start, *fib_ptr;
    const int2 *ent;           // nullptr => coordinate 0 (thread j owns line j)
    const int32_t *perm_in, *perm_out;
    int64_t stride;
    int32_t num_threads;
    int32_t n1;                // n + 1
    int32_t mode;
    int32_t fma;
};

__device__ __forceinline__ double mac_rt(double acc, double a, double b, int fma) {
    return fma ? __fma_rn(a, b, acc) : __dadd_rn(acc, __dmul_rn(a, b));
}

#if defined(LPFNT_MAXT) && LPFNT_MAXT == 8  // defined once, in the MAXT=8 translation unit
__global__ void __launch_bounds__(128) k_sweep_generic(const GenericArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.num_threads) return;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    int t, lane = 0, p0 = 0, base0 = 0;
    if (a.ent) {
        const int f = a.thr_fib[j];
        lane = j - a.thr_start[f];
        p0 = a.fib_ptr[f];
        const int nl = a.fib_ptr[f + 1] - p0;
        t = 0;
        while (t < nl && a.ent[p0 + t].y > lane) ++t;
    } else {
        base0 = a.line_off[j];
        t = a.line_off[j + 1] - base0;
    }
    auto pos = [&](int k) -> int { return a.ent ? a.ent[p0 + k].x + lane : base0 + k; };
    auto rd = [&](int k) -> double { const int e = pos(k); return in[a.perm_in ? a.perm_in[e] : e]; };
    auto wr = [&](int k, double v) { const int e = pos(k); out[a.perm_out ? a.perm_out[e] : e] = v; };
    const int n1 = a.n1;
    // NOTE: with a permutation the caller guarantees in != out, without one every thread only
    // touches its own fibre, so the in-place orders below are safe.
    if (a.mode == kLowerMatvec) {
        for (int k = t - 1; k >= 0; --k) {
            const double *row = a.mat + int64_t(k) * (k + 1) / 2;
            double acc = 0.0;
            for (int q = 0; q <= k; ++q) acc = mac_rt(acc, row[q], rd(q), a.fma);
            wr(k, acc);
        }
    } else if (a.mode == kUpperMatvec) {
        for (int k = 0; k < t; ++k) {
            const double *row = a.mat + int64_t(k) * n1 - int64_t(k) * (k - 1) / 2;
            double acc = 0.0;
            for (int q = k; q < t; ++q) acc = mac_rt(acc, row[q - k], rd(q), a.fma);
            wr(k, acc);
        }
    } else if (a.mode == kLowerSolve) {
        // needs out_j (j < k): read them back from `out` (written by this thread)
        for (int k = 0; k < t; ++k) {
            const double *row = a.mat + int64_t(k) * (k + 1) / 2;
            double s = 0.0;
            for (int q = 0; q < k; ++q) {
                const int e = pos(q);
                s = mac_rt(s, row[q], out[a.perm_out ? a.perm_out[e] : e], 0);
            }
            wr(k, __ddiv_rn(__dsub_rn(rd(k), s), row[k]));
        }
    } else {
        for (int k = t - 1; k >= 0; --k) {
            const double *row = a.mat + int64_t(k) * n1 - int64_t(k) * (k - 1) / 2;
            double s = mac_rt(0.0, row[0], 0.0, 0);
            for (int q = k + 1; q < t; ++q) {
                const int e = pos(q);
                s = mac_rt(s, row[q - k], out[a.perm_out ? a.perm_out[e] : e], 0);
            }
            wr(k, __ddiv_rn(__dsub_rn(rd(k), s), row[0]));
        }
    }
}

#endif

}  // namespace lpfnt
