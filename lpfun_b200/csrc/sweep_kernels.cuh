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

#ifndef LPFNT_GROUP_LANES
#define LPFNT_GROUP_LANES 8
#endif
constexpr int kRunLanes = LPFNT_GROUP_LANES;  // consecutive lanes of a fibre that stay together in k_sweep_groups (host: kGroupLanes)

enum SweepMode : int {
    kLowerMatvec = 0,  // out_k = sum_{j<=k} L[k,j] x_j                (fnt with inv V, ifnt, dxT)
    kUpperMatvec = 1,  // out_k = sum_{j>=k} U[k,j] x_j                (dx)
    kLowerSolve = 2,   // out_k = (x_k - sum_{j<k} L[k,j] out_j)/L[k,k] (fnt, precomputation=False)
    kUpperSolve = 3,   // out_k = (x_k - (U[k,k]*0 + sum_{j>k} U[k,j] out_j))/U[k,k], j ascending (coordinate 0:
                       // transform_ut_1d / the 1d stage of transform_ut_2d/3d/md, atoms.py:54-69, 372-385)
    kUpperSolveDesc = 4, // same, j DEscending: coordinates >= 1 accumulate the leaders last-to-first
                         // (atoms.py:391-402, 975-1003)
    kUpperMatvecDesc = 5 // out_k = sum_{j>=k} U[k,j] x_j folded from the LAST j down to k: dx on p = inf sets with m > 1, where the
                         // reference runs dtransform_max on reversed arrays (src/lpfun/core/molecules.py:290-294)
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
    } else if constexpr (MODE == kUpperMatvecDesc) {
        // row form, rows ascending (out_k only reads x_j, j >= k, so it may overwrite x_k), every row folded from j = t - 1 down
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            if (k < t) {
                double s = 0.0;
#pragma unroll
                for (int j = MAXT - 1; j >= k; --j)
                    if (j < t) s = mac<FMA>(s, M.v[j * (j + 1) / 2 + k], x[j]);
                x[k] = s;
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

template <int MAXT, int MODE, bool FMA, int ROWS = 4>
__device__ __forceinline__ void apply_fibre(double (&x)[MAXT], const int t, const PackedTri<MAXT> &M) {
    apply_fibre<MAXT, MODE, FMA, ROWS>(x, t, M, t);
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

// Colex permutation as a stand-alone streaming pass (used with the pipelined kernels, where it
// runs on an L2-resident chunk): gather y[e] = x[perm[e]]  /  scatter y[perm[e]] = x[e].
__global__ void __launch_bounds__(256) k_permute(const double *x, double *y, const int32_t *perm, int64_t n, int64_t stride, int scatter);

// Coarse <-> fine coefficient transfer through an ordinal embedding phi (Transform.embed, src/lpfun/transform.py:850-899,
// README "adaptivity" workflow): prolong fine[b][phi[e]] = coarse[b][e] (fine zero-filled before), restrict
// coarse[b][e] = fine[b][phi[e]].  Grid: (ceil(n_small / 256), batch).
__global__ void __launch_bounds__(256) k_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int restrict_);

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
    constexpr int GL = kRunLanes, NR = (CAP + GL - 1) / GL;
    int offs[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) offs[i] = (sub + GL * i < nl) ? __ldg(a.ent + p0 + sub + GL * i).x : 0;
    const int src_base = (threadIdx.x & 31) & ~(GL - 1);
    double x[CAP];
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        const int off = __shfl_sync(0xffffffffu, offs[k / GL], src_base + (k % GL));
        x[k] = (k < t) ? in[off + lane] : 0.0;
    }
    apply_fibre<CAP, MODE, FMA>(x, t, M, __reduce_max_sync(0xffffffffu, t));
    // the height is read again (L1 hit): otherwise the compiler keeps the "k < t" predicates of the
    // loads alive across the triangle code, packed into a bit mask (P2R/LOP3 chains, register pressure)
    const int t2 = live ? int(*reinterpret_cast<const volatile uint8_t *>(a.group_t + j)) : 0;
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        const int off = __shfl_sync(0xffffffffu, offs[k / GL], src_base + (k % GL));
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
    const int g = j / kRunLanes, sub = j % kRunLanes;
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
__global__ void __launch_bounds__(kSlabThreads, MAXT <= 24 ? 4 : 3) k_slab(const SlabArgs a, const __grid_constant__ PackedTri<MAXT> M) {
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

// Row blocks of a result pushed into the other ranks' copies of a symmetric buffer over NVLink (the all-gather of a batch
// sharded by rows as a plain streaming copy: 16-byte loads from local HBM, 16-byte stores into every peer mapping, or ONE
// multimem.st.v4 through the NVSwitch multicast mapping, which reaches every GPU).  Grid-stride over 16-byte words.
struct PushArgs {
    const uint4 *src;
    uint4 *dst[8];
    uint4 *mc;       // multicast mapping (or nullptr)
    int64_t words;   // 16-byte words
    int32_t n_dst;
};
__global__ void __launch_bounds__(128, 16) k_push_rows(const PushArgs a);

#if defined(LPFNT_MAXT) && LPFNT_MAXT == 8  // defined once, in the MAXT=8 translation unit
// 128 threads and <= 32 registers per thread: one such CTA fits next to the persistent sweep CTA of an SM (768 x 80 or
// 384 x 160 registers leave 4096), so the copy of chunk k can run while chunk k+1 is swept.
__global__ void __launch_bounds__(128, 16) k_push_rows(const PushArgs a) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < a.words; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(a.src + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (a.mc) {
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + i + u * stride), "f"(__uint_as_float(v[u].x)),
                             "f"(__uint_as_float(v[u].y)), "f"(__uint_as_float(v[u].z)), "f"(__uint_as_float(v[u].w))
                             : "memory");
            } else {
                for (int r = 0; r < a.n_dst; ++r) a.dst[r][i + u * stride] = v[u];
            }
        }
    }
    for (; i < a.words; i += stride) {
        const uint4 v = __ldcs(a.src + i);
        if (a.mc) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + i), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                         "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                         : "memory");
        } else {
            for (int r = 0; r < a.n_dst; ++r) a.dst[r][i] = v;
        }
    }
}
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
    } else if (a.mode == kUpperMatvecDesc) {
        for (int k = 0; k < t; ++k) {
            const double *row = a.mat + int64_t(k) * n1 - int64_t(k) * (k - 1) / 2;
            double acc = 0.0;
            for (int q = t - 1; q >= k; --q) acc = mac_rt(acc, row[q - k], rd(q), a.fma);
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
