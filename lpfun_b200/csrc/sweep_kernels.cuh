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
    kUpperSolve = 3,   // out_k = (x_k - (U[k,k]*0 + sum_{j>k} U[k,j] out_j))/U[k,k], j ascending (coordinate 0:
                       // transform_ut_1d / the 1d stage of transform_ut_2d/3d/md, atoms.py:54-69, 372-385)
    kUpperSolveDesc = 4  // same, j DEscending: coordinates >= 1 accumulate the leaders last-to-first
                         // (atoms.py:391-402, 975-1003)
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

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }


// Apply the triangular operator to a register-resident fibre of length t, in place.
// ROWS = output rows folded side by side in the lower mat-vec (independent accumulators).
// tg: bound used by the row-block guards of the lower mat-vec.  Callers may pass the WARP maximum of t
// (one value per warp: the guards then never diverge and the compiler does not have to keep 32 per-lane
// predicates alive); rows beyond a lane's own t are computed on zeros and never stored.
template <int MAXT, int MODE, bool FMA, int ROWS = 4>
__device__ __forceinline__ void apply_fibre(double (&x)[MAXT], const int t, const PackedTri<MAXT> &M, const int tg) {
    if constexpr (MODE == kLowerMatvec) {
        // Outputs from the top down: out_k only reads x_j, j <= k, so it may overwrite x_k.
        // ROWS rows are folded side by side (independent accumulators): the sweeps are latency
        // bound (ncu: "wait" is the top stall at 20 warps/SM), and several chains per thread hide the
        // 8.4-cycle DADD/DFMA latency; each row keeps its own ascending-j order.
        // Rows of a block that lie beyond the fibre (k >= t) are computed on zeros and never stored.
        static_assert(MAXT % ROWS == 0, "rows are processed in blocks of ROWS");
#pragma unroll
        for (int k = MAXT - 1; k >= ROWS - 1; k -= ROWS) {
            if (k - (ROWS - 1) < tg) {
                double acc[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) acc[r] = 0.0;
#pragma unroll
                for (int j = 0; j <= k - (ROWS - 1); ++j) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) acc[r] = mac<FMA>(acc[r], M.v[(k - r) * (k - r + 1) / 2 + j], x[j]);
                }
#pragma unroll
                for (int j = k - (ROWS - 1) + 1; j <= k; ++j) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r)
                        if (r <= k - j) acc[r] = mac<FMA>(acc[r], M.v[(k - r) * (k - r + 1) / 2 + j], x[j]);
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r) x[k - r] = acc[r];
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
    } else if constexpr (MODE == kUpperSolveDesc) {
#pragma unroll
        for (int k = MAXT - 1; k >= 0; --k) {
            if (k < t) {
                double s = 0.0;
#pragma unroll
                for (int j = MAXT - 1; j > k; --j)
                    if (j < t) s = mac<false>(s, M.v[j * (j + 1) / 2 + k], x[j]);
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

// Lower mat-vec with the INPUT STREAMED (shared memory) and the outputs accumulated in registers:
// step j reads x_j once and feeds it into out_j .. out_{tmax-1}.  Every out_k still receives its
// terms in ascending j from +0.0 (bit-identical to the row form), but a thread now carries up to
// MAXT independent chains instead of 2-4, so the 8.4-cycle DADD/DFMA latency is hidden inside ONE
// warp; the loads of x_j overlap the arithmetic of x_{j-1}.  tmax is warp-uniform (rows beyond it
// are skipped in blocks of four by uniform branches), t is the lane's own length (x_j = 0 beyond
// it; rows beyond it are computed on zeros and never stored).
template <int MAXT, bool FMA, typename LoadX>
__device__ __forceinline__ void lower_matvec_stream(double (&acc)[MAXT], const int tmax, const PackedTri<MAXT> &M, LoadX loadx) {
    static_assert(MAXT % 4 == 0, "rows are skipped in blocks of four");
#pragma unroll
    for (int k = 0; k < MAXT; ++k) acc[k] = 0.0;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
        if (j < tmax) {
            const double xj = loadx(j);
#pragma unroll
            for (int kb = j / 4; kb < MAXT / 4; ++kb) {
                if (kb * 4 < tmax) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int k = kb * 4 + r;
                        if (k >= j) acc[k] = mac<FMA>(acc[k], M.v[k * (k + 1) / 2 + j], xj);
                    }
                }
            }
        }
    }
}

template <int MAXT, int MODE, bool FMA, int ROWS = 4>
__device__ __forceinline__ void apply_fibre(double (&x)[MAXT], const int t, const PackedTri<MAXT> &M) {
    apply_fibre<MAXT, MODE, FMA, ROWS>(x, t, M, t);
}

// Two register-resident fibres swept side by side with ONE stream of matrix constants.
// Measured on B200 (tools/probe.cu): a warp that owns one fibre per lane is limited by the
// LDCU -> FP64 dependency (about 15 cycles per constant pair), far below the FP64 pipe; two fibres
// per lane halve the constants per term and double the independent chains.  `t` is the longer of
// the two lengths (the host pairs neighbours of the length-sorted order, so they differ little);
// the shorter fibre is zero padded, its surplus rows are computed but never stored.
template <int MAXT, int MODE, bool FMA>
__device__ __forceinline__ void apply_fibre2(double (&x)[MAXT], double (&y)[MAXT], const int t, const PackedTri<MAXT> &M) {
    if constexpr (MODE == kLowerMatvec) {
        static_assert(MAXT % 2 == 0, "rows are processed in pairs");
#pragma unroll
        for (int k = MAXT - 1; k >= 1; k -= 2) {
            if (k - 1 < t) {
                double a1 = 0.0, a0 = 0.0, b1 = 0.0, b0 = 0.0;
#pragma unroll
                for (int j = 0; j < k; ++j) {
                    const double m1 = M.v[k * (k + 1) / 2 + j], m0 = M.v[(k - 1) * k / 2 + j];
                    a1 = mac<FMA>(a1, m1, x[j]);
                    b1 = mac<FMA>(b1, m1, y[j]);
                    a0 = mac<FMA>(a0, m0, x[j]);
                    b0 = mac<FMA>(b0, m0, y[j]);
                }
                const double md = M.v[k * (k + 1) / 2 + k];
                x[k] = mac<FMA>(a1, md, x[k]);
                y[k] = mac<FMA>(b1, md, y[k]);
                x[k - 1] = a0;
                y[k - 1] = b0;
            }
        }
    } else if constexpr (MODE == kUpperMatvec) {
#pragma unroll
        for (int j = 0; j < MAXT; ++j) {
            if (j < t) {
                const double xj = x[j], yj = y[j];
#pragma unroll
                for (int k = 0; k < j; ++k) {
                    const double u = M.v[j * (j + 1) / 2 + k];
                    x[k] = mac<FMA>(x[k], u, xj);
                    y[k] = mac<FMA>(y[k], u, yj);
                }
                const double ud = M.v[j * (j + 1) / 2 + j];
                x[j] = mac<FMA>(0.0, ud, xj);
                y[j] = mac<FMA>(0.0, ud, yj);
            }
        }
    } else {
        apply_fibre<MAXT, MODE, FMA>(x, t, M);
        apply_fibre<MAXT, MODE, FMA>(y, t, M);
    }
}

struct LineSweepArgs {
    const double *in;
    double *out;
    const int32_t *line_off;  // N_1 + 1
    const int32_t *perm_in;   // gather: x[e] = in[perm_in[e]]  (or nullptr)
    const int32_t *perm_out;  // scatter: out[perm_out[e]] = y[e] (or nullptr)
    const int32_t *line_order; // lines ordered by length inside every chunk of blockDim.x lines
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

    // every element of the chunk is requested before anything waits (cp.async, no register staging)
    if (a.perm_in) {
        for (int ib = threadIdx.x; ib < cnt; ib += 8 * THREADS) {
            int src[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) src[u] = (ib + u * THREADS < cnt) ? __ldg(a.perm_in + e0 + ib + u * THREADS) : 0;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (ib + u * THREADS < cnt) cp_async8(buf + ib + u * THREADS, in + src[u]);
        }
    } else {
#pragma unroll 4
        for (int i = threadIdx.x; i < cnt; i += THREADS) cp_async8(buf + i, in + e0 + i);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    if (l0 + int(threadIdx.x) < l1) {
        const int l = a.line_order[l0 + threadIdx.x];  // the threadIdx.x-th longest line of the chunk
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
#pragma unroll 4
        for (int i = threadIdx.x; i < cnt; i += THREADS) out[__ldg(a.perm_out + e0 + i)] = buf[i];
    } else {
#pragma unroll 4
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
    const int4 *tile_rec;      // per tile {first row, #rows, first item, #items}
    const int2 *tile_elem;     // per tile {start in elem_row (multiple of 4), #packed elements}
    const int2 *ent;           // rows: {line element offset, line length}
    const uint16_t *row_pos;   // rows: packed shared-memory offset inside the tile
    const uint8_t *elem_row;   // packed elements: row (relative to the tile)
    const uint32_t *items;     // element-fibres, per tile sorted by length: row | lane << 16 | t << 24
    const uint32_t *rows0;     // rows, per tile sorted by length: row | len << 24
    const int32_t *perm_in;    // colex gather on load (or nullptr)
    const int32_t *perm_out;   // colex scatter on store (or nullptr)
    int64_t stride;
    int32_t debug_skip;        // development switch: 1 = skip the sweeps (times the data movement alone)
};

constexpr int kTileElemCap = 5120;  // packed elements per tile (host: kTileElems)

// Row-tile sweep along coordinate h >= 1, optionally fused with the coordinate-0 sweep.
//
// A tile is a run of consecutive fibres of coordinate h: every fibre is a (a_0, a_h) slice made of
// <= n+1 contiguous lines ("rows").  The rows of a tile are staged PACKED (back to back) in shared
// memory.  Everything is one-element-per-lane work:
//   copy in   packed element i comes from global  i + delta[row(i)]   (cp.async, all in flight at
//             once; row(i) from a uint8 map, delta[] = line offset - packed offset, in smem)
//   DIM0      thread <-> two rows (contiguous runs),          sorted by length
//   sweep h   thread <-> two element-fibres, element k at pos[row + k] + lane, sorted by length
//   copy out  the mirror image of copy in.
// ncu on the earlier row-per-warp copy showed 4.7 issued instructions per element at d=6 (rows of
// ~10 elements leave 2/3 of a warp idle in every copy instruction); here the copies cost ~0.3.
// The work items are ordered by length on the host, so the 32 lanes of a warp run triangles of
// nearly equal size, and each thread owns two NEIGHBOURS of that order (apply_fibre2).  Warp w of
// CTA b takes chunk (w + b) mod 4 of the sorted list.  Grid: (tiles, batch).
template <int MAXT, int MODE, bool FMA, bool DIM0, int THREADS>
__global__ void __launch_bounds__(THREADS, 3) k_sweep_tiles(const TileSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    __shared__ __align__(16) double buf[kTileElemCap];
    __shared__ int sdelta[2 * THREADS];
    __shared__ uint16_t spos[2 * THREADS];
    constexpr int NW = THREADS / 32;
    const int4 tr = __ldg(a.tile_rec + blockIdx.x);  // {first row, #rows, first item, #items}
    const int r0 = tr.x, nrows = tr.y, i0 = tr.z, nitems = tr.w;
    const int2 te = __ldg(a.tile_elem + blockIdx.x);
    const int nelem = te.y;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint8_t *erow = a.elem_row + te.x;
    for (int r = threadIdx.x; r < nrows; r += THREADS) {
        const int2 en = __ldg(a.ent + r0 + r);
        const int pos = __ldg(a.row_pos + r0 + r);
        spos[r] = static_cast<uint16_t>(pos);
        sdelta[r] = en.x - pos;
    }
    // sorted work-item descriptors of this thread; fetched now, used after the copy
    const int q = (((warp + blockIdx.x) % NW) << 5) + lane;
    uint32_t it0 = 0, it1 = 0, rw0 = 0, rw1 = 0;
    if (2 * q < nitems) it0 = __ldg(a.items + i0 + 2 * q);
    if (2 * q + 1 < nitems) it1 = __ldg(a.items + i0 + 2 * q + 1);
    if (DIM0 && 2 * q < nrows) rw0 = __ldg(a.rows0 + r0 + 2 * q);
    if (DIM0 && 2 * q + 1 < nrows) rw1 = __ldg(a.rows0 + r0 + 2 * q + 1);
    __syncthreads();

#pragma unroll 4
    for (int i = threadIdx.x; i < nelem; i += THREADS) {
        int src = i + sdelta[__ldg(erow + i)];
        if (a.perm_in) src = __ldg(a.perm_in + src);
        cp_async8(buf + i, in + src);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    if constexpr (DIM0) {
        if (2 * q < nrows) {
            const int t0 = int(rw0 >> 24), t1 = int(rw1 >> 24);
            double *p0 = buf + spos[rw0 & 0xffffu], *p1 = buf + spos[rw1 & 0xffffu];
            double x[MAXT], y[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) {
                x[k] = (k < t0) ? p0[k] : 0.0;
                y[k] = (k < t1) ? p1[k] : 0.0;
            }
            if (!a.debug_skip) apply_fibre2<MAXT, MODE, FMA>(x, y, t0, M);
#pragma unroll
            for (int k = 0; k < MAXT; ++k) {
                if (k < t0) p0[k] = x[k];
                if (k < t1) p1[k] = y[k];
            }
        }
        __syncthreads();
    }
    if (2 * q < nitems) {
        const int t0 = int(it0 >> 24), t1 = int(it1 >> 24);
        const uint16_t *s0 = spos + (it0 & 0xffffu), *s1 = spos + (it1 & 0xffffu);
        const int l0 = int((it0 >> 16) & 0xffu), l1 = int((it1 >> 16) & 0xffu);
        double x[MAXT], y[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            x[k] = (k < t0) ? buf[s0[k] + l0] : 0.0;
            y[k] = (k < t1) ? buf[s1[k] + l1] : 0.0;
        }
        if (!a.debug_skip) apply_fibre2<MAXT, MODE, FMA>(x, y, t0, M);
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            if (k < t0) buf[s0[k] + l0] = x[k];
            if (k < t1) buf[s1[k] + l1] = y[k];
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int i = threadIdx.x; i < nelem; i += THREADS) {
        int dst = i + sdelta[__ldg(erow + i)];
        if (a.perm_out) dst = __ldg(a.perm_out + dst);
        out[dst] = buf[i];
    }
}

// ---------------------------------------------------------------------------------------
// Persistent, software-pipelined row-tile sweep (the fast path).
//
// Same packed tiles and length-sorted work items as k_sweep_tiles, but a CTA walks a contiguous
// range of work units (function-major, tile-minor) and keeps three things in flight:
//   * the packed elements of unit u+1 travel global -> shared memory (cp.async, one element per
//     lane, every copy of the tile outstanding at once) into the second data buffer,
//   * the descriptors of unit u+2 (row table, element->row map, sorted items) travel into a
//     3-slot ring, so that no thread ever waits on a descriptor -> address -> data chain,
//   * the stores of unit u-1 drain.
// Measured motivation: with one tile per CTA lifetime the data movement ALONE ran at 1.7-2.9 TB/s
// (3 CTAs/SM, each a serial chain of 4-10 exposed global latencies); the sweeps themselves reach
// 55-67 % of the FP64 pipe only when 8 warps/SM are busy (tools/probe2.cu).  2 CTAs x 4 warps per
// SM, two fibres per thread.
// ---------------------------------------------------------------------------------------
struct PipeArgs {
    const double *in;
    double *out;
    const int4 *tile_rec;      // per tile {first row, #rows, first item, #items}
    const int2 *tile_elem;     // per tile {start in elem_row (multiple of 4), #packed elements}
    const int2 *row_desc;      // rows: {line element offset - packed offset, packed offset}
    const uint8_t *elem_row;   // packed elements: row (relative to the tile)
    const uint32_t *items;     // row | lane << 16 | t << 24, sorted per tile
    const uint32_t *rows0;     // row | len << 24, sorted per tile
    int64_t stride;            // elements per function
    int32_t num_tiles;
    int32_t batch;             // functions in this launch
    int32_t fuse0;             // 1: also sweep coordinate 0 (rows) before coordinate h
    int32_t units_per_cta;
    int32_t debug_skip;
};

constexpr int kPipeMaxUnits = 128;
constexpr int kPipeRows = 256;  // rows / items per tile (host: kTileRows, kTileItems)

// shared-memory carve-up of k_tiles_pipe (bytes)
constexpr size_t kPipeDataBytes = size_t(2) * kTileElemCap * sizeof(double);
constexpr size_t kPipeMapBytes = size_t(3) * kTileElemCap;
constexpr size_t kPipeRowBytes = size_t(3) * kPipeRows * sizeof(int2);
constexpr size_t kPipeItemBytes = size_t(3) * kPipeRows * sizeof(uint32_t);
constexpr size_t kPipeRecBytes = size_t(kPipeMaxUnits) * (sizeof(int4) + sizeof(int2));
constexpr size_t kPipeSmemBytes = kPipeDataBytes + kPipeMapBytes + kPipeRowBytes + 2 * kPipeItemBytes + kPipeRecBytes;

template <int MAXT, int MODE, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_tiles_pipe(const PipeArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *data = reinterpret_cast<double *>(smem_raw);                                        // [2][cap]
    uint8_t *smap = smem_raw + kPipeDataBytes;                                                  // [3][cap]
    int2 *srow = reinterpret_cast<int2 *>(smap + kPipeMapBytes);                                // [3][256] {delta, pos}
    uint32_t *sitem = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(srow) + kPipeRowBytes);  // [3][256]
    uint32_t *srow0 = sitem + 3 * kPipeRows;                                                    // [3][256]
    int4 *strec = reinterpret_cast<int4 *>(srow0 + 3 * kPipeRows);                              // [units]
    int2 *stel = reinterpret_cast<int2 *>(strec + kPipeMaxUnits);                               // [units]

    const int64_t total = int64_t(a.num_tiles) * a.batch;
    const int64_t ubeg = int64_t(blockIdx.x) * a.units_per_cta;
    const int64_t left = total - ubeg;
    const int nunits = left < a.units_per_cta ? int(left) : a.units_per_cta;
    if (nunits <= 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = THREADS / 32;

    for (int u = tid; u < nunits; u += THREADS) {
        const int tile = int((ubeg + u) % a.num_tiles);
        strec[u] = __ldg(a.tile_rec + tile);
        stel[u] = __ldg(a.tile_elem + tile);
    }
    __syncthreads();

    auto issue_desc = [&](int u) {  // descriptors of local unit u -> ring slot u % 3
        const int4 tr = strec[u];
        const int2 te = stel[u];
        const int slot = u % 3;
        for (int r = tid; r < tr.y; r += THREADS) {
            cp_async8(srow + slot * kPipeRows + r, a.row_desc + tr.x + r);
            if (a.fuse0) cp_async4(srow0 + slot * kPipeRows + r, a.rows0 + tr.x + r);
        }
        for (int j = tid; j < tr.w; j += THREADS) cp_async4(sitem + slot * kPipeRows + j, a.items + tr.z + j);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.elem_row + te.x);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smap + slot * kTileElemCap);
        for (int j = tid; 4 * j < te.y; j += THREADS) cp_async4(dst + j, src + j);
    };
    auto issue_data = [&](int u) {  // packed elements of local unit u -> data buffer u % 2 (needs its descriptors)
        const int nelem = stel[u].y;
        const int2 *rows = srow + (u % 3) * kPipeRows;
        const uint8_t *map = smap + (u % 3) * kTileElemCap;
        double *dst = data + (u & 1) * kTileElemCap;
        const double *src = a.in + ((ubeg + u) / a.num_tiles) * a.stride;
#pragma unroll 4
        for (int i = tid; i < nelem; i += THREADS) cp_async8(dst + i, src + (i + rows[map[i]].x));
    };

    // prologue: descriptors 0 and 1, then the data of unit 0
    issue_desc(0);
    if (nunits > 1) issue_desc(1);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    issue_data(0);
    cp_async_commit();

    for (int u = 0; u < nunits; ++u) {
        cp_async_wait_all();
        __syncthreads();  // data(u), descriptors(u+1) landed; every thread is done with unit u-1
        if (u + 1 < nunits) issue_data(u + 1);
        if (u + 2 < nunits) issue_desc(u + 2);
        cp_async_commit();

        const int4 tr = strec[u];
        const int nelem = stel[u].y;
        const int slot = u % 3;
        double *cur = data + (u & 1) * kTileElemCap;
        const int2 *rows = srow + slot * kPipeRows;
        // chunk of the sorted lists for this warp (rotated per unit: long chunks visit every SMSP)
        const int q = (((warp + u + blockIdx.x) % NW) << 5) + lane;
        if (a.fuse0) {
            if (2 * q < tr.y) {
                const uint32_t d0 = srow0[slot * kPipeRows + 2 * q];
                const uint32_t d1 = (2 * q + 1 < tr.y) ? srow0[slot * kPipeRows + 2 * q + 1] : 0u;
                const int t0 = int(d0 >> 24), t1 = int(d1 >> 24);
                double *p0 = cur + rows[d0 & 0xffffu].y, *p1 = cur + rows[d1 & 0xffffu].y;
                double x[MAXT], y[MAXT];
#pragma unroll
                for (int k = 0; k < MAXT; ++k) {
                    x[k] = (k < t0) ? p0[k] : 0.0;
                    y[k] = (k < t1) ? p1[k] : 0.0;
                }
                if (!a.debug_skip) apply_fibre2<MAXT, MODE, FMA>(x, y, t0, M);
#pragma unroll
                for (int k = 0; k < MAXT; ++k) {
                    if (k < t0) p0[k] = x[k];
                    if (k < t1) p1[k] = y[k];
                }
            }
            __syncthreads();
        }
        if (2 * q < tr.w) {
            const uint32_t d0 = sitem[slot * kPipeRows + 2 * q];
            const uint32_t d1 = (2 * q + 1 < tr.w) ? sitem[slot * kPipeRows + 2 * q + 1] : 0u;
            const int t0 = int(d0 >> 24), t1 = int(d1 >> 24);
            const int2 *s0 = rows + (d0 & 0xffffu), *s1 = rows + (d1 & 0xffffu);
            const int l0 = int((d0 >> 16) & 0xffu), l1 = int((d1 >> 16) & 0xffu);
            double x[MAXT], y[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) {
                x[k] = (k < t0) ? cur[s0[k].y + l0] : 0.0;
                y[k] = (k < t1) ? cur[s1[k].y + l1] : 0.0;
            }
            if (!a.debug_skip) apply_fibre2<MAXT, MODE, FMA>(x, y, t0, M);
#pragma unroll
            for (int k = 0; k < MAXT; ++k) {
                if (k < t0) cur[s0[k].y + l0] = x[k];
                if (k < t1) cur[s1[k].y + l1] = y[k];
            }
        }
        __syncthreads();
        {
            const uint8_t *map = smap + slot * kTileElemCap;
            double *dst = a.out + ((ubeg + u) / a.num_tiles) * a.stride;
#pragma unroll 4
            for (int i = tid; i < nelem; i += THREADS) dst[i + rows[map[i]].x] = cur[i];
        }
    }
}

// Colex permutation as a stand-alone streaming pass (used with the pipelined kernels, where it
// runs on an L2-resident chunk): gather y[e] = x[perm[e]]  /  scatter y[perm[e]] = x[e].
__global__ void __launch_bounds__(256) k_permute(const double *x, double *y, const int32_t *perm, int64_t n, int64_t stride, int scatter);

// Coarse <-> fine coefficient transfer through an ordinal embedding phi (Transform.embed, src/lpfun/transform.py:850-899,
// README "adaptivity" workflow): prolong fine[b][phi[e]] = coarse[b][e] (fine zero-filled before), restrict
// coarse[b][e] = fine[b][phi[e]].  Grid: (ceil(n_small / 256), batch).
__global__ void __launch_bounds__(256) k_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int restrict_);

// ---------------------------------------------------------------------------------------
// Length-sorted sweep straight from global memory: thread j owns the j-th LONGEST element-fibre
// of coordinate h (h = 0: the j-th longest line).  No shared memory, no barriers: the 32 lanes of
// a warp run triangles of equal size (the natural lane <-> a_0 mapping keeps only 17 % (d=6) to
// 50 % (d=3, n=30) of the FP64 lanes busy), at the price of less coalesced 8-byte accesses that
// the L1 absorbs (neighbours of the natural order stay neighbours inside a length class).
// Grid: (ceil(items / THREADS), batch).
// ---------------------------------------------------------------------------------------
struct SortedSweepArgs {
    const double *in;
    double *out;
    const int2 *sorted;       // {first ent entry (h >= 1) | line offset (h = 0), lane | t << 8}
    const int2 *ent;          // rows of coordinate h (unused for h = 0)
    const int32_t *perm_in;
    const int32_t *perm_out;
    int64_t stride;
    int32_t num_items;
};

template <int MAXT, int MODE, bool FMA, bool DIM0, int THREADS>
__global__ void __launch_bounds__(THREADS) k_sweep_sorted(const SortedSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    const int j = blockIdx.x * THREADS + threadIdx.x;
    if (j >= a.num_items) return;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    const int2 it = __ldg(a.sorted + j);
    const int t = it.y >> 8, lane = it.y & 0xff;
    double x[MAXT];
    if constexpr (DIM0) {
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            x[k] = 0.0;
            if (k < t) x[k] = in[a.perm_in ? a.perm_in[it.x + k] : it.x + k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            x[k] = 0.0;
            if (k < t) x[k] = in[__ldg(a.ent + it.x + k).x + lane];
        }
    }
    apply_fibre<MAXT, MODE, FMA>(x, t, M);
    if constexpr (DIM0) {
#pragma unroll
        for (int k = 0; k < MAXT; ++k)
            if (k < t) out[a.perm_out ? a.perm_out[it.x + k] : it.x + k] = x[k];
    } else {
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            if (k < t) {
                const int e = __ldg(a.ent + it.x + k).x + lane;
                out[a.perm_out ? a.perm_out[e] : e] = x[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Coordinate h >= 1 with height-ordered lane groups: like k_sweep_fibres, but the unit that
// stays together is a run of <= 8 consecutive lanes of one fibre (64 contiguous bytes per row),
// and the runs are ordered by height.  A warp holds four runs of similar height, so the unrolled
// triangle code is not dragged to the tallest fibre of an arbitrary neighbourhood (the natural
// order keeps 17 % of the FP64 lanes busy at d=6, n=20).  No shared memory, no barriers.
// Grid: (ceil(8 * groups / THREADS), batch).
// ---------------------------------------------------------------------------------------
struct GroupSweepArgs {
    const double *in;
    double *out;
    const int2 *groups;       // {first ent entry, lane0 | height << 8 | lanes << 16}
    const uint8_t *group_t;   // 8 per group: height of every lane
    const int2 *ent;
    const int32_t *perm_out;
    int64_t stride;
    int32_t num_groups;
    int32_t two_lanes;        // 1: k_sweep_groups2 (two adjacent lanes per thread)
    const uint8_t *cta_cap;   // per CTA of 128 threads (16 runs): height of its tallest fibre (or nullptr)
};

// Body of k_sweep_groups for a row capacity CAP <= MAXT: the packed triangle of capacity MAXT starts with the one of
// capacity CAP (row-major lower form), so the same kernel parameter serves every capacity.
template <int CAP, int MAXT, int MODE, bool FMA>
__device__ __forceinline__ void sweep_group_body(const GroupSweepArgs &a, const PackedTri<MAXT> &Mfull, const int j, const bool live,
                                                 const int p0, const int lane, const int nl, const int sub, const int t,
                                                 const double *in, double *out) {
    const PackedTri<CAP> &M = reinterpret_cast<const PackedTri<CAP> &>(Mfull);
    // The <= CAP row offsets of the run are fetched ONCE, up to four per thread (rows sub, sub+8, ...),
    // and handed around the 8 threads of the run with shuffles: one global latency ahead of the
    // data loads instead of a dependent descriptor load per row, and nothing to re-load for the stores
    // (ncu: descriptor look-ups and their dependent compares were 2/3 of the issued instructions).
    constexpr int NR = CAP / 8;
    int offs[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) offs[i] = (sub + 8 * i < nl) ? __ldg(a.ent + p0 + sub + 8 * i).x : 0;
    const int src_base = (threadIdx.x & 31) & ~7;
    double x[CAP];
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        const int off = __shfl_sync(0xffffffffu, offs[k >> 3], src_base + (k & 7));
        x[k] = (k < t) ? in[off + lane] : 0.0;
    }
    apply_fibre<CAP, MODE, FMA>(x, t, M, __reduce_max_sync(0xffffffffu, t));
    // the height is read again (L1 hit): otherwise the compiler keeps the "k < t" predicates of the
    // loads alive across the triangle code, packed into a bit mask (P2R/LOP3 chains, register pressure)
    const int t2 = live ? int(*reinterpret_cast<const volatile uint8_t *>(a.group_t + j)) : 0;
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        const int off = __shfl_sync(0xffffffffu, offs[k >> 3], src_base + (k & 7));
        if (k < t2) {
            const int e = off + lane;
            out[a.perm_out ? a.perm_out[e] : e] = x[k];
        }
    }
}

// 80 registers (6 CTAs/SM) for 32-long fibres, 64 registers (8 CTAs/SM) up to 24 rows: measured -6..10 % per pass at d=6, n=20.
// The runs are ordered by height, so whole CTAs late in the grid only hold short fibres: cta_cap[blockIdx.x] (host) is the
// tallest fibre of the CTA, and the CTA runs the body unrolled for 8, 16 or MAXT rows -- in high dimensions most fibres are
// short (d=6, n=20: 6 rows on average at a capacity of 24) and the never-taken rows' shuffle / address / predicate
// instructions were most of the issued instructions (ncu: 70 % issue-slot utilisation).
template <int MAXT, int MODE, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS, MAXT <= 24 ? 8 : 6) k_sweep_groups(const GroupSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    const int j = blockIdx.x * THREADS + threadIdx.x;
    const int g = j >> 3, sub = j & 7;
    const bool live = g < a.num_groups;  // whole warps stay alive for the shuffles below
    const int2 gr = live ? __ldg(a.groups + g) : make_int2(0, 0);
    const int p0 = gr.x, lane = (gr.y & 0xff) + sub, nl = (gr.y >> 8) & 0xff;
    const int t = live ? int(__ldg(a.group_t + j)) : 0;  // height of this lane (0: lane absent)
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    // (measured: helps at capacities 16 / 24 -- d=6 n=20: 34 -> 30 us per pass --, hurts by 6 % at 32, where the runs are
    // long and the batched launches keep every SM full of tall CTAs anyway)
    const int cap = (MAXT <= 24 && a.cta_cap) ? int(__ldg(a.cta_cap + blockIdx.x)) : MAXT;
    if constexpr (MAXT > 8 && MAXT <= 24) {
        if (cap <= 8) { sweep_group_body<8, MAXT, MODE, FMA>(a, M, j, live, p0, lane, nl, sub, t, in, out); return; }
    }
    if constexpr (MAXT > 16 && MAXT <= 24) {
        if (cap <= 16) { sweep_group_body<16, MAXT, MODE, FMA>(a, M, j, live, p0, lane, nl, sub, t, in, out); return; }
    }
    sweep_group_body<MAXT, MAXT, MODE, FMA>(a, M, j, live, p0, lane, nl, sub, t, in, out);
}

// ---------------------------------------------------------------------------------------
// Whole-function kernel for m <= 3: ONE pass over HBM per transform.
//
// A function of an m <= 3 dimensional space with N <= ~28 000 coefficients (d=3, n=30: 15 216 =
// 119 KB) fits in the shared memory of one SM.  A persistent CTA walks functions b = blockIdx.x,
// blockIdx.x + gridDim.x, ...; per function it
//   1. copies the N values in (cp.async, 16 B per lane; 8 B per lane through the colex permutation),
//   2. sweeps coordinate 0, 1, (2) IN shared memory -- thread <-> element-fibre(s) in registers, the
//      items of every coordinate ordered by length (host) so that a warp's lanes run equal
//      triangles, chunks rotated over the warps per function --,
//   3. copies the N results out (coalesced; 8 B scatter through the permutation for ifnt).
// All descriptors (sorted items, per-coordinate row offsets) are loaded ONCE per CTA lifetime, so a
// function costs one exposed global latency; 16 B/DOF of HBM traffic per transform instead of
// m x 16 B/DOF.  SMs de-synchronise, so some are always in their copy phase while others compute.
// ITEMS = fibres per thread and round (2 shares the matrix constants between two fibres: FMA form).
// ---------------------------------------------------------------------------------------
struct WholeArgs {
    const double *in;
    double *out;
    const int2 *sorted[3];     // per coordinate: items {first entry in ent_off (h >= 1) | line offset (h = 0), lane | t << 8}
    const uint16_t *ent_off[3]; // per coordinate h >= 1: line offsets in fibre order (N_1 entries); [0] unused
    const int32_t *perm_in;    // colex gather on the way in (or nullptr)
    const int32_t *perm_out;   // colex scatter on the way out (or nullptr)
    int32_t n_elem;            // N
    int32_t n_lines;           // N_1 = items per coordinate
    int32_t m;                 // 1..3
    int32_t batch;
    long long *dbg;            // development: clock64 stamps of CTA 0 (k_whole2; nullptr in production)
};

template <int MAXT, int MODE, bool FMA, int ITEMS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_whole(const WholeArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.n_elem, N1 = a.n_lines, m = a.m;
    const int Npad = (N + 1) & ~1;
    double *buf = reinterpret_cast<double *>(smem_raw);                    // [Npad]
    int2 *sitem = reinterpret_cast<int2 *>(buf + Npad);                    // [3][N1]: {eo index | t << 24, element offset added}
    uint16_t *seo = reinterpret_cast<uint16_t *>(sitem + 3 * N1);          // [32 identity][N1 coordinate 1][N1 coordinate 2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = THREADS / 32;

    // One addressing form for every coordinate: element k of an item sits at seo[p + k] + add.
    //   coordinate 0 (lines):   p = 0 (identity table), add = line offset
    //   coordinate h >= 1:      p = 32 + (h-1) N1 + first entry of the fibre, add = lane
    if (tid < 32) seo[tid] = static_cast<uint16_t>(tid);
    for (int h = 0; h < m; ++h) {
        for (int i = tid; i < N1; i += THREADS) {
            const int2 it = __ldg(a.sorted[h] + i);
            const int t = it.y >> 8;
            if (h == 0) sitem[i] = make_int2(t << 24, it.x);
            else {
                sitem[h * N1 + i] = make_int2((32 + (h - 1) * N1 + it.x) | (t << 24), it.y & 0xff);
                seo[32 + (h - 1) * N1 + i] = __ldg(a.ent_off[h] + i);
            }
        }
    }
    const int rounds = (N1 + THREADS * ITEMS - 1) / (THREADS * ITEMS);

    for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        const double *in = a.in + int64_t(b) * N;
        double *out = a.out + int64_t(b) * N;
        __syncthreads();  // previous function fully stored (and the tables above written)
        if (a.perm_in) {
#pragma unroll 4
            for (int i = tid; i < N; i += THREADS) cp_async8(buf + i, in + __ldg(a.perm_in + i));
        } else if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
            for (int i = 2 * tid; i + 1 < N; i += 2 * THREADS) cp_async16(buf + i, in + i);
            if ((N & 1) && tid == 0) cp_async8(buf + N - 1, in + N - 1);
        } else {
#pragma unroll 4
            for (int i = tid; i < N; i += THREADS) cp_async8(buf + i, in + i);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();

#pragma unroll 1
        for (int h = 0; h < m; ++h) {
            const int2 *items = sitem + h * N1;
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) {
                // slot of this thread in the sorted list: chunks rotate over the warps from function to
                // function so that the long chunks visit every SM sub-partition
                // (snake order over the rounds: a warp that drew a long chunk in one round draws a short
                // one in the next, which keeps the barrier tails short)
                const int wrot = (warp + b) % NW;
                const int chunk = r * NW + ((r & 1) ? (NW - 1 - wrot) : wrot);
                if constexpr (ITEMS == 1) {
                    const int q = chunk * 32 + lane;
                    int t = 0, add = 0;
                    const uint16_t *eo = seo;
                    if (q < N1) { const int2 it = items[q]; t = int(unsigned(it.x) >> 24); eo = seo + (it.x & 0xffffff); add = it.y; }
                    double x[MAXT];
#pragma unroll
                    for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[eo[k] + add] : 0.0;
                    apply_fibre<MAXT, MODE, FMA>(x, t, M);
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) buf[eo[k] + add] = x[k];
                } else {
                    const int q = chunk * 64 + 2 * lane;
                    int t0 = 0, a0 = 0, t1 = 0, a1 = 0;
                    const uint16_t *e0 = seo, *e1 = seo;
                    if (q < N1) { const int2 it = items[q]; t0 = int(unsigned(it.x) >> 24); e0 = seo + (it.x & 0xffffff); a0 = it.y; }
                    if (q + 1 < N1) { const int2 it = items[q + 1]; t1 = int(unsigned(it.x) >> 24); e1 = seo + (it.x & 0xffffff); a1 = it.y; }
                    double x[MAXT], y[MAXT];
#pragma unroll
                    for (int k = 0; k < MAXT; ++k) {
                        x[k] = (k < t0) ? buf[e0[k] + a0] : 0.0;
                        y[k] = (k < t1) ? buf[e1[k] + a1] : 0.0;
                    }
                    apply_fibre2<MAXT, MODE, FMA>(x, y, t0, M);
#pragma unroll
                    for (int k = 0; k < MAXT; ++k) {
                        if (k < t0) buf[e0[k] + a0] = x[k];
                        if (k < t1) buf[e1[k] + a1] = y[k];
                    }
                }
            }
            __syncthreads();
        }
        if (a.perm_out) {
#pragma unroll 4
            for (int i = tid; i < N; i += THREADS) out[__ldg(a.perm_out + i)] = buf[i];
        } else if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            for (int i = 2 * tid; i + 1 < N; i += 2 * THREADS) *reinterpret_cast<double2 *>(out + i) = *reinterpret_cast<const double2 *>(buf + i);
            if ((N & 1) && tid == 0) out[N - 1] = buf[N - 1];
        } else {
#pragma unroll 4
            for (int i = tid; i < N; i += THREADS) out[i] = buf[i];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Whole-function kernel, second generation (m = 2, 3; N_1 <= THREADS): ONE pass over HBM per transform.
//
//   * persistent: one CTA per SM, 768 threads; the CTA walks over its share of the functions;
//   * the function (N doubles, contiguous in HBM) arrives in shared memory by the bulk-copy
//     engine (cp.async.bulk + mbarrier), no thread touches a load instruction for it;
//   * every sweep is ONE round: thread <-> item (line for coordinate 0, element-fibre otherwise),
//     items in length order so that warps are uniform; coordinates 0 .. m-2 work in place in
//     shared memory;
//   * the LAST coordinate reads its fibres into registers, then the buffer is free: one thread
//     issues the bulk copy of the CTA's NEXT function while everybody computes the last sweep
//     and stores the result from registers straight to HBM (fused colex scatter).  The load of
//     function k+1 therefore overlaps the last third of the arithmetic of function k, and the
//     FP64 pipe only idles at the four barriers per function.
// Algorithmic traffic = actual traffic = 16 B/DOF per transform.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gsrc), "r"(bytes), "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar)))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_src))), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kWhole2Threads = 768;
constexpr unsigned kBulkPiece = 32768;  // bytes per bulk-copy instruction

template <int MAXT, int MODE, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_whole2(const WholeArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = a.n_elem, N1 = a.n_lines, m = a.m;
    const int Npad = (N + 3) & ~1;  // N + 1 (odd N: the bulk copy moves an even number of elements) + 1 (shifted origin)
    double *buf0 = reinterpret_cast<double *>(smem_raw);                   // [Npad], 16-byte aligned landing zone
    uint64_t *bar = reinterpret_cast<uint64_t *>(buf0 + Npad);             // one mbarrier
    int2 *sitem = reinterpret_cast<int2 *>(bar + 2);                       // [m][N1]: {eo index | t << 24, element offset added}
    uint16_t *seo = reinterpret_cast<uint16_t *>(sitem + 3 * N1);          // [32 identity][N1 coordinate 1][N1 coordinate 2]
    uint16_t *sperm = seo + ((32 + 2 * N1 + 7) & ~7);                      // [N] colex permutation (N < 65536), if any
    const int tid = threadIdx.x;
    {   // the permutation is looked up once per element and function: from shared memory (30-cycle latency) instead of
        // L1/L2 (measured: the dependent global look-ups made the scatter stores 37 % of an ifnt function's period)
        const int32_t *pg = a.perm_in ? a.perm_in : a.perm_out;
        if (pg)
            for (int e = tid; e < N; e += THREADS) sperm[e] = static_cast<uint16_t>(__ldg(pg + e));
    }

    // One addressing form for every coordinate: element k of an item sits at seo[p + k] + add (see k_whole).
    if (tid < 32) seo[tid] = static_cast<uint16_t>(tid);
    for (int h = 0; h < m; ++h) {
        for (int i = tid; i < N1; i += THREADS) {
            const int2 it = __ldg(a.sorted[h] + i);
            const int t = it.y >> 8;
            if (h == 0) sitem[i] = make_int2(t << 24, it.x);
            else {
                sitem[h * N1 + i] = make_int2((32 + (h - 1) * N1 + it.x) | (t << 24), it.y & 0xff);
                seo[32 + (h - 1) * N1 + i] = __ldg(a.ent_off[h] + i);
            }
        }
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Bulk copies move 16-byte granules between 16-byte aligned addresses.  For odd N every second function starts 8
    // bytes off: the copy then starts one element early (mis = 1) and the function sits at buf0 + mis; an even-b function
    // of an odd N copies one element more than it needs (the next function's first) -- except the very last function of
    // the batch, which copies N - 1 elements and fetches its last one with an ordinary load (nothing is read past the end).
    const bool oddN = (N & 1) != 0;
    auto issue_load = [&](int b) {
        const int mis = oddN ? (b & 1) : 0;
        const bool tail = oddN && mis == 0 && b == a.batch - 1;
        const unsigned nbytes = unsigned(oddN ? (tail ? N - 1 : N + 1) : N) * 8u;
        const char *src = reinterpret_cast<const char *>(a.in + int64_t(b) * N - mis);
        mbar_expect_tx(bar, nbytes);
        for (unsigned o = 0; o < nbytes; o += kBulkPiece)
            bulk_g2s(reinterpret_cast<char *>(buf0) + o, src + o, min(kBulkPiece, nbytes - o), bar);
        if (tail) buf0[N - 1] = a.in[int64_t(b) * N + N - 1];  // ordered before the consumers by the barrier after the mbarrier wait
    };
    if (tid == 0 && int(blockIdx.x) < a.batch) issue_load(blockIdx.x);
    unsigned parity = 0;

    // Item slot of this thread: the chunks of 32 items (longest first) are dealt to the warps in SNAKE order over
    // the four SM sub-partitions (warp w runs on sub-partition w % 4): 0 1 2 3 / 3 2 1 0 / ...  With the plain
    // order sub-partition 0 would get the longest chunk of every round (13 % more FP64 work than the average).
    const int slot = ((tid >> 7) * 4 + (((tid >> 7) & 1) ? 3 - ((tid >> 5) & 3) : ((tid >> 5) & 3))) * 32 + (tid & 31);
    // development timeline: [function 0..3][warp][stamp 0..15] clock64 values of CTA 0
    int fn = 0;
    auto stamp = [&](int id) {
        if (a.dbg && blockIdx.x == 0 && fn < 4 && (tid & 31) == 0) {
            long long c;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
            a.dbg[(fn * (THREADS / 32) + (tid >> 5)) * 16 + id] = c;
        }
    };
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x, ++fn) {
        double *out = a.out + int64_t(b) * N;
        double *buf = buf0 + (oddN ? (b & 1) : 0);
        stamp(0);
        mbar_wait(bar, parity);
        parity ^= 1;
        if (oddN && b == a.batch - 1) __syncthreads();  // the last function's final element came by an ordinary store of thread 0
        stamp(1);
#pragma unroll 1
        for (int h = 0; h < m; ++h) {
            int t = 0, add = 0;
            const uint16_t *eo = seo;
            if (slot < N1) { const int2 it = sitem[h * N1 + slot]; t = int(unsigned(it.x) >> 24); eo = seo + (it.x & 0xffffff); add = it.y; }
            const bool last = (h == m - 1);
            double x[MAXT];
            if (h == 0 && a.perm_in) {
#pragma unroll
                for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[sperm[add + k]] : 0.0;
                __syncthreads();  // the gather reads other lines' slots: nobody writes before everybody has read
            } else if (h == 0) {
                // coordinate 0: the line is contiguous, no offset look-ups
#pragma unroll
                for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[add + k] : 0.0;
            } else {
#pragma unroll
                for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[eo[k] + add] : 0.0;
            }
            if (last) {
                // every thread holds its fibre in registers: the buffer is free for the next function
                fence_proxy_async();
                __syncthreads();
                if (tid == 0 && b + int(gridDim.x) < a.batch) issue_load(b + gridDim.x);
            }
            stamp(2 + 4 * h);  // loads issued (and, where a barrier follows them, complete)
            // two rows side by side: with 24 resident warps that is enough to hide the FP64 latency, and it leaves the
            // register allocator room (no spills); measured 0.643 / 0.488 ms (fnt / ifnt) against 0.672 / 0.496 with four
            // rows, 0.731 / 0.543 with eight, 0.656 / 0.531 with one
            apply_fibre<MAXT, MODE, FMA, 2>(x, t, M, __reduce_max_sync(0xffffffffu, t));
            stamp(3 + 4 * h);  // triangle code issued
            // the item is read again: nothing but the fibre stays in registers across the triangle code
            // (in particular not the 32 "k < t" predicates, which the compiler would pack into a bit mask)
            if (slot < N1) {
                const volatile int *vi = reinterpret_cast<const volatile int *>(sitem + h * N1 + slot);
                const int w0 = vi[0];
                t = int(unsigned(w0) >> 24); eo = seo + (w0 & 0xffffff); add = vi[1];
            }
            if (last) {
                if (a.perm_out) {
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) out[sperm[eo[k] + add]] = x[k];
                } else {
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) out[eo[k] + add] = x[k];
                }
                stamp(4 + 4 * h);  // stores issued
            } else {
                if (h == 0) {
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) buf[add + k] = x[k];
                } else {
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) buf[eo[k] + add] = x[k];
                }
                stamp(4 + 4 * h);  // stores issued
                __syncthreads();
                stamp(5 + 4 * h);  // barrier passed
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Whole-function kernel, third generation ("box" layout; m = 2, 3): ONE pass over HBM per
// transform, no offset look-ups, producer warps.
//
// Shared-memory layout of a function: plane k (= alpha_2; for m = 2 every line is its own plane)
// is a rectangle of nl[k] rows with an ODD row stride S[k] >= longest line of the plane:
//       addr(a0, a1, k) = P[k] + a1 * S[k] + a0           (P, S: kernel constants, BoxConst)
// so a fibre along coordinate 1 is a constant-stride walk, a fibre along coordinate 2 needs only
// the constants, and the lines of one plane (coordinate 0, thread <-> line) fall into different
// banks.  ~4/pi of the compact size for an l^2 ball (152 KB at d=3, n=30).
//
// Thread roles: 736 compute threads (23 warps), thread <-> one item per sweep, items in length
// order (uniform warps); 2 producer warps copy the CTA's NEXT function from HBM into the box
// (cp.async 8 B, colex gather applied on the way) while the compute warps run the last sweep
// from registers and store its result straight to HBM (colex scatter applied on the way).
// Named barriers: FULL (producers -> compute), EMPTY (compute -> producers), STEP (compute only).
// ---------------------------------------------------------------------------------------
struct BoxConst {
    int32_t P[32];  // first element of plane k in the box
    int32_t S[32];  // row stride of plane k
    int32_t L[32];  // first line of plane k in the compact (HBM) order
};

struct Whole3Args {
    const double *in;
    double *out;
    const uint32_t *item0;   // lines:               box base | t << 16
    const uint32_t *item1;   // (a0, k) fibres:      box base | stride << 16 | t << 24        (m = 3 only)
    const uint32_t *item2;   // (a0, a1) fibres:     a0 | a1 << 8 | t << 16
    const int2 *pline;       // per line (compact order): {box base | t << 16, compact offset}
    const uint16_t *perm16;  // colex permutation (or nullptr)
    const uint16_t *line_off16;  // compact offset of every line
    int32_t gather, scatter;
    int32_t n_elem, n_box, n_lines, m, batch;
    int32_t n_item0, n_item1, n_item2;
};

constexpr int kW3Compute = 736;
constexpr int kW3Producer = 32;
inline size_t whole3_smem_bytes(int n_box, int n_lines, int n_elem, bool perm) {
    return size_t(n_box) * 8 + size_t(3) * kW3Compute * 4 + size_t(n_lines) * 8 + size_t((n_lines + 1) & ~1) * 2 + (perm ? size_t(n_elem + 8) * 2 : 0) + 16;
}
constexpr int kW3Threads = kW3Compute + kW3Producer;

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int MAXT, int MODE, bool FMA>
__global__ void __launch_bounds__(kW3Threads, 1)
k_whole3(const Whole3Args a, const __grid_constant__ PackedTri<MAXT> M, const __grid_constant__ BoxConst C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *buf = reinterpret_cast<double *>(smem_raw);                 // [n_box]
    unsigned *sitem = reinterpret_cast<unsigned *>(buf + a.n_box);      // [3][kW3Compute] item words (0 = idle)
    int2 *spline = reinterpret_cast<int2 *>(sitem + 3 * kW3Compute);    // [n_lines] producer's line descriptors
    uint16_t *soff = reinterpret_cast<uint16_t *>(spline + a.n_lines);  // [n_lines (+1 if odd)]
    uint16_t *sperm = soff + ((a.n_lines + 1) & ~1);                    // [n_elem] colex permutation (if any)
    const int tid = threadIdx.x;
    const int N = a.n_elem;
    constexpr int FULL = 1, EMPTY = 2, STEP = 3;

    for (int i = tid; i < a.n_lines; i += kW3Threads) {
        soff[i] = __ldg(a.line_off16 + i);
        spline[i] = __ldg(a.pline + i);
    }
    if (a.perm16)
        for (int i = tid; i < N; i += kW3Threads) sperm[i] = __ldg(a.perm16 + i);
    if (tid < kW3Compute) {
        sitem[tid] = tid < a.n_item0 ? __ldg(a.item0 + tid) : 0u;
        sitem[kW3Compute + tid] = (a.m == 3 && tid < a.n_item1) ? __ldg(a.item1 + tid) : 0u;
        sitem[2 * kW3Compute + tid] = tid < a.n_item2 ? __ldg(a.item2 + tid) : 0u;
    }
    __syncthreads();

    if (tid >= kW3Compute) {
        // ---------------- producer warp: one coalesced cp.async per line ----------------
        const int lane = tid - kW3Compute;
        bool first = true;
        for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
            if (!first) named_sync(EMPTY, kW3Threads);  // the compute warps hold function b - grid in registers
            first = false;
            const double *in = a.in + int64_t(b) * N;
#pragma unroll 4
            for (int l = 0; l < a.n_lines; ++l) {
                const int2 d = spline[l];  // {box base | t << 16, compact offset}
                if (lane < (d.x >> 16)) {
                    int src = d.y + lane;
                    if (a.gather) src = sperm[src];
                    cp_async8(buf + (d.x & 0xffff) + lane, in + src);
                }
            }
            cp_async_commit();
            cp_async_wait_all();
            named_arrive(FULL, kW3Threads);
        }
        return;
    }

    // ---------------- compute warps ----------------
    // The item words live in shared memory and are read again after the triangle code: nothing but
    // the fibre itself stays in registers across apply_fibre (80 registers per thread at 768 threads).
    const volatile unsigned *my_item = sitem + tid;
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        named_sync(FULL, kW3Threads);
#pragma unroll 1
        for (int h = 0; h < 3; ++h) {
            if (h == 1 && a.m == 2) continue;
            double x[MAXT];
            int t;
            {
                const unsigned it = my_item[h * kW3Compute];
                if (h < 2) {
                    // coordinate 0: thread <-> line (stride 1); coordinate 1: thread <-> (a0, k), constant stride
                    t = (h == 0) ? int(it >> 16) : int(it >> 24);
                    const int st = (h == 0) ? 1 : int((it >> 16) & 0xff);
                    const double *q = buf + (it & 0xffff);
#pragma unroll
                    for (int k = 0; k < MAXT; ++k) {
                        x[k] = (k < t) ? *q : 0.0;
                        q += st;
                    }
                } else {
                    // last coordinate: thread <-> (a0, a1); planes through the constants
                    t = int(it >> 16);
                    const int a0 = int(it & 0xff), a1 = int((it >> 8) & 0xff);
#pragma unroll
                    for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[C.P[k] + a1 * C.S[k] + a0] : 0.0;
                    // the function now lives in registers: hand the box to the producers
                    if (b + int(gridDim.x) < a.batch) named_arrive(EMPTY, kW3Threads);
                }
            }
            apply_fibre<MAXT, MODE, FMA, FMA ? 4 : 2>(x, t, M);
            const unsigned it = my_item[h * kW3Compute];
            if (h < 2) {
                const int st = (h == 0) ? 1 : int((it >> 16) & 0xff);
                double *p = buf + (it & 0xffff);
#pragma unroll
                for (int k = 0; k < MAXT; ++k) {
                    if (k < t) *p = x[k];
                    p += st;
                }
                named_sync(STEP, kW3Compute);
            } else {
                double *out = a.out + int64_t(b) * N;
                const int a0 = int(it & 0xff), a1 = int((it >> 8) & 0xff);
                if (a.scatter) {
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) out[sperm[int(soff[C.L[k] + a1]) + a0]] = x[k];
                } else {
#pragma unroll
                    for (int k = 0; k < MAXT; ++k)
                        if (k < t) out[int(soff[C.L[k] + a1]) + a0] = x[k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Slab kernel (any m >= 2, any batch): coordinates 0 AND 1 in ONE pass over HBM.
//
// The sweeps along coordinates 0 and 1 only couple elements of the same 2-D plane
// (a_2 .. a_{m-1} fixed).  A slab is a run of consecutive planes (<= kSlabThreads lines,
// <= kSlabThreads element-fibres of coordinate 1, <= kSlabElems elements: contiguous in HBM).
// One CTA works on one (function, slab) unit at a time:
//   cp.async the slab into shared memory (fused colex gather) -> coordinate-0 sweep, thread <->
//   line, lines in length order, in place -> barrier -> coordinate-1 sweep, thread <-> element-
//   fibre in length order, read from shared memory into registers -> result stored from
//   registers straight to HBM (fused colex scatter when coordinate 1 is the last one, m = 2).
// Small footprint (<= 54 KB shared memory, 80 registers x 256 threads): three CTAs per SM run
// unsynchronised, so the copy / barrier phases of one hide behind the arithmetic of the others --
// the property the whole-function kernels (one CTA per SM) lack.  Units are slab-major, so the
// CTAs running at any one time share their item tables through L2.
// ---------------------------------------------------------------------------------------
constexpr int kSlabThreads = 256;
constexpr int kSlabElems = 5440;  // 43.5 KB + tables < 48 KB static shared memory

struct SlabDesc {
    int32_t e0, ne;      // first element, number of elements
    int32_t l0, nl;      // first line, number of lines
    int32_t c0, nc;      // first coordinate-1 item, number of them
    int32_t pad0, pad1;
};

struct SlabArgs {
    const double *in;
    double *out;
    const SlabDesc *slab;
    const int2 *item0;        // per slab (from l0): lines, longest first: {t << 24, offset relative to the slab}
    const int2 *item1;        // per slab (from c0): fibres, longest first: {32 + first line relative to the slab | t << 24, lane}
    const int32_t *line_off;  // N_1 + 1
    const int32_t *perm_in;   // colex gather (or nullptr)
    const int32_t *perm_out;  // colex scatter (or nullptr; only when coordinate 1 is the last sweep)
    int64_t stride;           // N
    int32_t n_slabs, batch;
};

template <int MAXT, int MODE, bool FMA>
__global__ void __launch_bounds__(kSlabThreads, 3) k_slab(const SlabArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    __shared__ __align__(16) double buf[kSlabElems];
    __shared__ int2 sitem0[kSlabThreads], sitem1[kSlabThreads];
    __shared__ uint16_t seo[32 + kSlabThreads];
    const int tid = threadIdx.x;
    const int64_t units = int64_t(a.n_slabs) * a.batch;
    int cur = -1;
    SlabDesc d{};
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int s = int(u / a.batch), b = int(u - int64_t(s) * a.batch);
        const double *in = a.in + int64_t(b) * a.stride;
        double *out = a.out + int64_t(b) * a.stride;
        __syncthreads();  // the previous unit has taken everything out of shared memory
        if (s != cur) {
            cur = s;
            d = a.slab[s];
            sitem0[tid] = tid < d.nl ? __ldg(a.item0 + d.l0 + tid) : make_int2(0, 0);
            sitem1[tid] = tid < d.nc ? __ldg(a.item1 + d.c0 + tid) : make_int2(0, 0);
            if (tid < 32) seo[tid] = uint16_t(tid);
            if (tid < d.nl) seo[32 + tid] = uint16_t(__ldg(a.line_off + d.l0 + tid) - d.e0);
        }
        if (a.perm_in) {
            // all gather indices first (independent loads in flight), then the copies
            constexpr int U = (kSlabElems + kSlabThreads - 1) / kSlabThreads;
            int src[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = tid + u * kSlabThreads;
                src[u] = (i < d.ne) ? __ldg(a.perm_in + d.e0 + i) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (src[u] >= 0) cp_async8(buf + tid + u * kSlabThreads, in + src[u]);
        } else {
#pragma unroll 4
            for (int i = tid; i < d.ne; i += kSlabThreads) cp_async8(buf + i, in + d.e0 + i);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        {   // coordinate 0
            const int2 it = sitem0[tid];
            const int t = int(unsigned(it.x) >> 24);
            const double *p = buf + it.y;
            double x[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? p[k] : 0.0;
            apply_fibre<MAXT, MODE, FMA, 4>(x, t, M, __reduce_max_sync(0xffffffffu, t));
            // the item is read again: nothing but the fibre stays in registers across the triangle code
            // (in particular not the 32 "k < t" predicates, which the compiler would pack into a bit mask)
            const volatile int *vi = reinterpret_cast<volatile int *>(sitem0 + tid);
            const int t2 = int(unsigned(vi[0]) >> 24);
            double *q = buf + vi[1];
#pragma unroll
            for (int k = 0; k < MAXT; ++k)
                if (k < t2) q[k] = x[k];
        }
        __syncthreads();
        {   // coordinate 1: element k of the fibre sits at seo[p + k] + lane
            const int2 it = sitem1[tid];
            const int t = int(unsigned(it.x) >> 24);
            const uint16_t *eo = seo + (it.x & 0xffffff);
            double x[MAXT];
#pragma unroll
            for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? buf[eo[k] + it.y] : 0.0;
            apply_fibre<MAXT, MODE, FMA, 4>(x, t, M, __reduce_max_sync(0xffffffffu, t));
            // read the item again: nothing but the fibre stays in registers across the triangle code
            const volatile int *vi = reinterpret_cast<volatile int *>(sitem1 + tid);
            const int2 it2 = make_int2(vi[0], vi[1]);
            const int t2 = int(unsigned(it2.x) >> 24);
            const uint16_t *eo2 = seo + (it2.x & 0xffffff);
            double *o = out + d.e0 + it2.y;
            if (a.perm_out) {
#pragma unroll
                for (int k = 0; k < MAXT; ++k)
                    if (k < t2) out[__ldg(a.perm_out + d.e0 + it2.y + eo2[k])] = x[k];
            } else {
#pragma unroll
                for (int k = 0; k < MAXT; ++k)
                    if (k < t2) o[eo2[k]] = x[k];
            }
        }
    }
}

// Two ADJACENT lanes per thread (same fibre, lanes 2s and 2s+1 of a height-ordered lane group):
// the row descriptor, the address arithmetic and the matrix constants are shared by two element-
// fibres, which halves the non-FP64 instructions per term (ncu: 2.5-3.8 issued instructions per
// term besides the DMUL/DADD in the one-lane kernels).  Heights never increase along the lanes,
// so the first lane's height bounds both (apply_fibre2).
template <int MAXT, int MODE, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS, 3) k_sweep_groups2(const GroupSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    const int j = blockIdx.x * THREADS + threadIdx.x;
    const int g = j >> 2, sub = (j & 3) * 2;
    if (g >= a.num_groups) return;
    const int2 gr = __ldg(a.groups + g);
    const int lanes = (gr.y >> 16) & 0xff;
    if (sub >= lanes) return;
    const bool two = sub + 1 < lanes;
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    const int p0 = gr.x, lane = (gr.y & 0xff) + sub, nl = (gr.y >> 8) & 0xff;
    double x[MAXT], y[MAXT];
    int t0 = 0, t1 = 0;
#pragma unroll
    for (int k = 0; k < MAXT; ++k) {
        x[k] = 0.0;
        y[k] = 0.0;
        if (k < nl) {
            const int2 en = __ldg(a.ent + p0 + k);
            if (en.y > lane) {
                const double *src = in + en.x + lane;
                x[k] = src[0];
                t0 = k + 1;
                if (two && en.y > lane + 1) {
                    y[k] = src[1];
                    t1 = k + 1;
                }
            }
        }
    }
    apply_fibre2<MAXT, MODE, FMA>(x, y, t0, M);
#pragma unroll
    for (int k = 0; k < MAXT; ++k) {
        if (k < t0) {
            const int e = __ldg(a.ent + p0 + k).x + lane;
            if (a.perm_out) {
                out[a.perm_out[e]] = x[k];
                if (k < t1) out[a.perm_out[e + 1]] = y[k];
            } else {
                out[e] = x[k];
                if (k < t1) out[e + 1] = y[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Coordinates 0 AND 1 in one pass, one WARP per plane (fixed a_2.., all a_0, a_1).
// A plane is <= n+1 contiguous lines of <= n+1 elements.  The warp copies the lines into its
// private slab of shared memory (row pitch MAXT+1, coalesced loads, descriptors handed around
// with shuffles), sweeps coordinate 0 with lane <-> line, re-reads the slab transposed
// (lane <-> a_0) for coordinate 1 and stores straight from registers (lanes = consecutive a_0:
// coalesced).  No CTA barrier (only __syncwarp), four shared-memory trips per element, one pass
// over memory fewer than the per-coordinate kernels.  The lanes of a warp run triangles of
// different size here (heights fall along a plane), i.e. some FP64 lanes idle -- the sweeps
// are latency / pass-count bound, not FP64 bound, so the saved pass wins.
// Grid: (ceil(planes / 4), batch), 128 threads.
// ---------------------------------------------------------------------------------------
struct PlaneSweepArgs {
    const double *in;
    double *out;
    const int32_t *line_off;   // N_1 + 1
    const int32_t *plane_line; // planes + 1: first line of every plane (CSR)
    const int32_t *perm_in;
    const int32_t *perm_out;
    int64_t stride;
    int32_t num_planes;
};

template <int MAXT, int MODE, bool FMA>
__global__ void __launch_bounds__(128, 6) k_sweep_planes(const PlaneSweepArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    constexpr int PAD = MAXT + 1;
    __shared__ double slab[4][MAXT * PAD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane = blockIdx.x * 4 + warp;
    if (plane >= a.num_planes) return;
    double *buf = slab[warp];
    const double *in = a.in + int64_t(blockIdx.y) * a.stride;
    double *out = a.out + int64_t(blockIdx.y) * a.stride;
    const int l0 = __ldg(a.plane_line + plane), nl = __ldg(a.plane_line + plane + 1) - l0;
    // lane r owns the descriptor of row r
    int my_off = 0, my_len = 0;
    if (lane < nl) {
        my_off = __ldg(a.line_off + l0 + lane);
        my_len = __ldg(a.line_off + l0 + lane + 1) - my_off;
    }
    // copy in: row r by all lanes (lane = a_0)
    // (cp.async: every row of the plane is in flight at once; with the colex gather the permutation
    // entries are fetched eight rows at a time first)
    if (a.perm_in) {
        for (int rb = 0; rb < nl; rb += 8) {
            int src[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int off = __shfl_sync(0xffffffffu, my_off, (rb + u) & 31);
                const int len = __shfl_sync(0xffffffffu, my_len, (rb + u) & 31);
                src[u] = (rb + u < nl && lane < len) ? __ldg(a.perm_in + off + lane) : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (src[u] >= 0) cp_async8(buf + (rb + u) * PAD + lane, in + src[u]);
        }
    } else {
#pragma unroll 4
        for (int r = 0; r < nl; ++r) {
            const int off = __shfl_sync(0xffffffffu, my_off, r);
            const int len = __shfl_sync(0xffffffffu, my_len, r);
            if (lane < len) cp_async8(buf + r * PAD + lane, in + off + lane);
        }
    }
    cp_async_commit();
    // height of column a_0 = lane: rows whose length exceeds it (lengths never increase along a_1)
    int t1 = 0;
#pragma unroll 4
    for (int r = 0; r < nl; ++r) t1 += (__shfl_sync(0xffffffffu, my_len, r) > lane) ? 1 : 0;
    cp_async_wait_all();
    __syncwarp();
    {   // coordinate 0: lane <-> line
        const int t = my_len;
        double *row = buf + lane * PAD;
        double x[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? row[k] : 0.0;
        apply_fibre<MAXT, MODE, FMA>(x, t, M);
#pragma unroll
        for (int k = 0; k < MAXT; ++k)
            if (k < t) row[k] = x[k];
    }
    __syncwarp();
    {   // coordinate 1: lane <-> a_0, element k = row k
        double x[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) x[k] = (k < t1) ? buf[k * PAD + lane] : 0.0;
        apply_fibre<MAXT, MODE, FMA>(x, t1, M);
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            const int off = __shfl_sync(0xffffffffu, my_off, k);
            if (k < t1) {
                const int e = off + lane;
                out[a.perm_out ? a.perm_out[e] : e] = x[k];
            }
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
__global__ void __launch_bounds__(256) k_permute(const double *x, double *y, const int32_t *perm, int64_t n, int64_t stride, int scatter) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double *xf = x + int64_t(blockIdx.y) * stride;
    double *yf = y + int64_t(blockIdx.y) * stride;
    const int64_t q = perm[e];
    if (scatter) yf[q] = xf[e];
    else yf[e] = xf[q];
}

__global__ void __launch_bounds__(256) k_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int restrict_) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_small) return;
    const int64_t q = phi[e];
    if (restrict_) out[int64_t(blockIdx.y) * n_small + e] = in[int64_t(blockIdx.y) * n_big + q];
    else out[int64_t(blockIdx.y) * n_big + q] = in[int64_t(blockIdx.y) * n_small + e];
}

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
    } else if (a.mode == kUpperSolveDesc) {
        for (int k = t - 1; k >= 0; --k) {
            const double *row = a.mat + int64_t(k) * n1 - int64_t(k) * (k - 1) / 2;
            double s = 0.0;
            for (int q = t - 1; q > k; --q) {
                const int e = pos(q);
                s = mac_rt(s, row[q - k], out[a.perm_out ? a.perm_out[e] : e], 0);
            }
            wr(k, __ddiv_rn(__dsub_rn(rd(k), s), row[0]));
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
