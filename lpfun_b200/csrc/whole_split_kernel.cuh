// Split whole-function kernel (m = 3, lower mat-vec): one pass over HBM per transform with TWO CTAs per SM.
//
// Reference: itransform_lt_3d (src/lpfun/core/atoms.py:668-742), i.e. the three sweeps of a batched fnt / ifnt
// (src/lpfun/transform.py:461-475, 581-595) fused over shared memory -- like k_whole (whole_kernel.cuh).
//
// Why: a d=3, n=30 function (119 KB) fills the shared memory of an SM once, so k_whole runs ONE CTA per SM and all its warps
// load, fold and store in lock step between barriers: the FP64 pipe idles during the load / store phases, the LSU during the
// triangle phases (phase timeline: 28 K of 43 K cycles are arithmetic for fnt, 15 K of 34 K for ifnt).  Here every function is
// processed as two units that fit twice into an SM, so that the phases of two independent CTAs overlap:
//   unit P = planes a_2 <  C (contiguous in memory),   unit Q = planes a_2 >= C.
// Coordinates 0 and 1 never couple different planes: they run inside each unit, in place in shared memory.  Last coordinate:
//   unit P: rows k < C of every column (a_0, a_1) -- a lower mat-vec on the fibre cut at C;
//   unit Q: rows k >= C of the columns that reach beyond C, every row folded from its FIRST term: the y[j], j < C, are
//           unit P's coordinate-1 results, exported by unit P to a per-CTA scratch [a_2][column] in global memory (L2
//           resident, 56 KB per CTA) and read back, coalesced, by unit Q; the y[j], j >= C, come from Q's own buffer.
// A CTA handles P and Q of one function back to back (no synchronisation between CTAs).  Every output receives exactly the
// reference's sequence of operations (ascending j from +0.0; EXACT: multiply and add rounded separately), the FP64 work is
// partitioned, not repeated; the extra traffic is the scratch (C / (n+1) of the columns' data, written and read once, in L2).
//
// Fill: bulk copy (cp.async.bulk + mbarrier) of the unit's contiguous element range, issued by one thread as soon as the
// previous unit has its last fibres in registers; with the colex gather (fnt) every thread fetches its share through the
// permutation with cp.async once its own fibre is dead.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "whole_kernel.cuh"

namespace lpfnt {

constexpr int kSplitThreads = 384;
constexpr int kSplitStamps = 32;  // clock64 stamps per warp and function (development timeline): 13 per unit

struct WholeSplitArgs {
    const double *in;
    double *out;
    const uint32_t *items;   // item lists of part P (coordinates 0, 1, 2) and part Q, back to back, in thread order
    const uint16_t *seo;     // row offsets inside the part, every fibre's list on a multiple of four entries
    const uint16_t *perm16;  // colex permutation (or nullptr)
    double *scratch;         // gridDim.x * C * ncolq doubles
    long long *dbg;          // development: clock64 stamps (nullptr in production)
    int32_t item_off[2][3], n_items[2][3];
    int32_t e0[2], ne[2];
    int32_t cb[32], rq[32];  // column (a_0, a_1) of Q = cb[a_1] + a_0 for a_0 < rq[a_1]
    int32_t total_items, seo_len, ncolq, C;
    int32_t buf_elems;       // doubles of shared memory for the data (even, >= max(ne) + 4)
    int32_t gather, scatter;
    int32_t n_elem, batch;
    int32_t skew;            // cycles the second CTA of every SM waits before its first unit (de-phases the two CTAs)
};

inline size_t whole_split_smem_bytes(int buf_elems, int total_items, int seo_len, int n_elem, bool perm) {
    return size_t(buf_elems) * 8 + 16 + size_t(total_items) * 4 + size_t((seo_len + 7) & ~7) * 2 + (perm ? size_t((n_elem + 7) & ~7) * 2 : 0);
}

// lower mat-vec, two rows side by side, rows below row_lo skipped (a pair that straddles row_lo is computed as a whole)
template <int MAXT, bool FMA>
__device__ __forceinline__ void apply_fibre_from(double (&x)[MAXT], const PackedTri<MAXT> &M, const int tg, const int row_lo) {
#pragma unroll
    for (int k = MAXT - 1; k >= 1; k -= 2) {
        if (k - 1 < tg && k >= row_lo) {
            double a1 = 0.0, a0 = 0.0;
#pragma unroll
            for (int j = 0; j < k; ++j) {
                a1 = mac<FMA>(a1, M.v[k * (k + 1) / 2 + j], x[j]);
                a0 = mac<FMA>(a0, M.v[(k - 1) * k / 2 + j], x[j]);
            }
            a1 = mac<FMA>(a1, M.v[k * (k + 1) / 2 + k], x[k]);
            x[k] = a1;
            x[k - 1] = a0;
        }
    }
}

template <int MAXT, bool FMA, bool TL = false>
__global__ void __launch_bounds__(kSplitThreads, 2) k_whole_split(const __grid_constant__ WholeSplitArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    constexpr int TG = kSplitThreads;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *const buf0 = reinterpret_cast<double *>(smem_raw);
#define S_BAR (reinterpret_cast<uint64_t *>(buf0 + a.buf_elems))
#define S_ITEM (reinterpret_cast<uint32_t *>(buf0 + a.buf_elems + 2))
#define S_SEO (reinterpret_cast<uint16_t *>(S_ITEM + a.total_items))
#define S_PERM (S_SEO + ((a.seo_len + 7) & ~7))
#define S_M(r0) (*reinterpret_cast<const PackedTri<MAXT> *>(&M.v[((r0) >> 20) << 1]))  // see whole_kernel.cuh: no hoisting, no spills
    const int N = a.n_elem;
    for (int i = threadIdx.x; i < a.total_items; i += TG) S_ITEM[i] = __ldg(a.items + i);
    for (int i = threadIdx.x; i < a.seo_len; i += TG) S_SEO[i] = __ldg(a.seo + i);
    if (a.gather || a.scatter)
        for (int e = threadIdx.x; e < N; e += TG) S_PERM[e] = __ldg(a.perm16 + e);
    if (threadIdx.x == 0) {
        mbar_init(S_BAR, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double *const scratch = a.scratch + size_t(blockIdx.x) * size_t(a.C) * size_t(a.ncolq);
    const int64_t total_elems = int64_t(a.batch) * N;

    // unit (b, p): elements [b N + e0[p], + ne[p]) -> buf0 + mis.  Bulk copies move 16-byte granules between 16-byte aligned
    // addresses: a range that starts on an odd element is copied from one element earlier (mis = 1), a range with an odd
    // count takes one element more -- except at the very end of the batch, where the last element comes by an ordinary load.
    auto unit_mis = [&](int b, int p) { return int((int64_t(b) * N + a.e0[p]) & 1); };
    auto issue_bulk = [&](int b, int p) {  // one thread
        const int64_t g = int64_t(b) * N + a.e0[p];
        const int mis = int(g & 1);
        int cnt = a.ne[p] + mis;
        bool tail = false;
        if (cnt & 1) {
            if (g - mis + cnt + 1 <= total_elems) cnt += 1;
            else { cnt -= 1; tail = true; }
        }
        const unsigned nbytes = unsigned(cnt) * 8u;
        const char *src = reinterpret_cast<const char *>(a.in + g - mis);
        mbar_expect_tx(S_BAR, nbytes);
        for (unsigned o = 0; o < nbytes; o += kBulkPiece)
            bulk_g2s(reinterpret_cast<char *>(buf0) + o, src + o, min(kBulkPiece, nbytes - o), S_BAR);
        if (tail) buf0[mis + a.ne[p] - 1] = a.in[g + a.ne[p] - 1];
    };
    auto issue_gather = [&](int b, int p) {  // every thread its share
        const double *src = a.in + int64_t(b) * N;
        const uint16_t *pm = S_PERM + a.e0[p];
        const int ne = a.ne[p];
#pragma unroll 4
        for (int i = threadIdx.x; i < ne; i += TG) cp_async8(buf0 + i, src + pm[i]);
        cp_async_commit();
    };

    int fn = 0;
    auto stamp = [&](int id) {
        if constexpr (TL) {
            if (a.dbg && blockIdx.x < 2 && fn < 4 && (threadIdx.x & 31) == 0) {
                long long c;
                asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
                a.dbg[((fn * 2 + blockIdx.x) * (TG / 32) + (threadIdx.x >> 5)) * kSplitStamps + id] = c;
            }
        }
    };

    if (int(blockIdx.x) < a.batch) {
        if (a.gather) issue_gather(blockIdx.x, 0);
        else if (threadIdx.x == 0) issue_bulk(blockIdx.x, 0);
    }
    // Two CTAs that start together on an SM run the same phases at the same time and gain nothing from sharing it: the
    // CTAs of the second wave start `skew` cycles late.
    if (a.skew > 0 && 2 * int(blockIdx.x) >= int(gridDim.x)) {
        long long c0, c1;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(c0)::"memory");
        do { asm volatile("mov.u64 %0, %%clock64;" : "=l"(c1)::"memory"); } while (c1 - c0 < a.skew);
    }
    unsigned parity = 0;
#pragma unroll 1
    for (int u = 0;; ++u) {
        const int b = int(blockIdx.x) + (u >> 1) * int(gridDim.x), p = u & 1;
        if (b >= a.batch) break;
        const int nb = p ? b + int(gridDim.x) : b, np = p ^ 1;  // the CTA's next unit
        const bool have_next = nb < a.batch;
        const int sb = p * 13;
        stamp(sb + 0);
        if (a.gather) {
            cp_async_wait_all();
            __syncthreads();
        } else {
            mbar_wait(S_BAR, parity);
            parity ^= 1;
            // the tail element of the very last range is an ordinary store of thread 0
            if (int64_t(b) * N + a.e0[p] + a.ne[p] == total_elems) __syncthreads();
        }
        stamp(sb + 1);
        double *const buf = buf0 + (a.gather ? 0 : unit_mis(b, p));
        const int C = a.C;
#pragma unroll 1
        for (int h = 0; h < 3; ++h) {
            const bool last = (h == 2);
#pragma unroll 1
            for (int r0 = 0; r0 < a.n_items[p][h]; r0 += TG) {
                const bool last_round = last && r0 + TG >= a.n_items[p][h];
                double x[MAXT];
                int t;
                {
                    const uint32_t it = S_ITEM[a.item_off[p][h] + r0 + threadIdx.x];
                    t = int((it >> 16) & 63u);
                    if (h == 0) {
                        const double *q = buf + (it & 0xffffu);
#pragma unroll
                        for (int k = 0; k < MAXT; ++k) x[k] = (k < t) ? q[k] : 0.0;
                    } else {
                        const uint2 *eo4 = reinterpret_cast<const uint2 *>(S_SEO + (it & 0xffffu));
                        const double *q = buf + ((it >> 22) & 31u);
                        // unit Q, last coordinate: the list starts with C placeholders; rows j < C come from the scratch
                        const int k_lo = (last && p == 1) ? C : 0;
#pragma unroll
                        for (int k4 = 0; k4 < MAXT; k4 += 4) {
                            const uint2 e = eo4[k4 >> 2];
                            x[k4] = (k4 >= k_lo && k4 < t) ? q[e.x & 0xffffu] : 0.0;
                            x[k4 + 1] = (k4 + 1 >= k_lo && k4 + 1 < t) ? q[e.x >> 16] : 0.0;
                            x[k4 + 2] = (k4 + 2 >= k_lo && k4 + 2 < t) ? q[e.y & 0xffffu] : 0.0;
                            x[k4 + 3] = (k4 + 3 >= k_lo && k4 + 3 < t) ? q[e.y >> 16] : 0.0;
                        }
                        if (last && p == 1 && t > 0) {
                            const double *sc = scratch + a.cb[(it >> 27) & 31u] + ((it >> 22) & 31u);
#pragma unroll
                            for (int k = 0; k < 16 && k < MAXT; ++k)
                                if (k < C) x[k] = __ldcg(sc + k * a.ncolq);
                        }
                    }
                }
                if (last_round) {
                    // every thread holds its fibre in registers: the buffer is free for the CTA's next unit
                    fence_proxy_async();
                    __syncthreads();
                    if (!a.gather && threadIdx.x == 0 && have_next) issue_bulk(nb, np);
                }
                stamp(sb + 2 + 4 * h);
                const int tg = __reduce_max_sync(0xffffffffu, t);
                if (last && p == 1) apply_fibre_from<MAXT, FMA>(x, S_M(r0), tg, C);
                else apply_fibre<MAXT, kLowerMatvec, FMA, 2>(x, t, S_M(r0), tg);
                stamp(sb + 3 + 4 * h);
                {   // the item is read again: nothing but the fibre stays in registers across the triangle code
                    const uint32_t it = *reinterpret_cast<const volatile uint32_t *>(S_ITEM + a.item_off[p][h] + r0 + threadIdx.x);
                    const int t2 = int((it >> 16) & 63u);
                    if (last) {
                        const uint2 *eo4 = reinterpret_cast<const uint2 *>(S_SEO + (it & 0xffffu));
                        const unsigned add = unsigned(a.e0[p]) + ((it >> 22) & 31u);
                        const int k_lo = p ? C : 0;
                        double *out = a.out + int64_t(b) * N;
                        if (a.scatter) {
#pragma unroll
                            for (int k8 = 0; k8 < MAXT; k8 += 8) {
                                const uint2 e0 = eo4[k8 >> 2], e1 = eo4[(k8 >> 2) + 1];
                                const unsigned o[8] = {e0.x & 0xffffu, e0.x >> 16, e0.y & 0xffffu, e0.y >> 16,
                                                       e1.x & 0xffffu, e1.x >> 16, e1.y & 0xffffu, e1.y >> 16};
                                unsigned uu[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) uu[j] = (k8 + j >= k_lo && k8 + j < t2) ? S_PERM[o[j] + add] : 0u;
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (k8 + j >= k_lo && k8 + j < t2) out[uu[j]] = x[k8 + j];
                            }
                        } else {
#pragma unroll
                            for (int k4 = 0; k4 < MAXT; k4 += 4) {
                                const uint2 e = eo4[k4 >> 2];
                                if (k4 >= k_lo && k4 < t2) out[(e.x & 0xffffu) + add] = x[k4];
                                if (k4 + 1 >= k_lo && k4 + 1 < t2) out[(e.x >> 16) + add] = x[k4 + 1];
                                if (k4 + 2 >= k_lo && k4 + 2 < t2) out[(e.y & 0xffffu) + add] = x[k4 + 2];
                                if (k4 + 3 >= k_lo && k4 + 3 < t2) out[(e.y >> 16) + add] = x[k4 + 3];
                            }
                        }
                    } else if (h == 0) {
                        double *q = buf + (it & 0xffffu);
#pragma unroll
                        for (int k = 0; k < MAXT; ++k)
                            if (k < t2) q[k] = x[k];
                    } else {
                        const uint2 *eo4 = reinterpret_cast<const uint2 *>(S_SEO + (it & 0xffffu));
                        const unsigned lane = (it >> 22) & 31u;
                        double *q = buf + lane;
#pragma unroll
                        for (int k4 = 0; k4 < MAXT; k4 += 4) {
                            const uint2 e = eo4[k4 >> 2];
                            if (k4 < t2) q[e.x & 0xffffu] = x[k4];
                            if (k4 + 1 < t2) q[e.x >> 16] = x[k4 + 1];
                            if (k4 + 2 < t2) q[e.y & 0xffffu] = x[k4 + 2];
                            if (k4 + 3 < t2) q[e.y >> 16] = x[k4 + 3];
                        }
                        if (p == 0) {
                            // export for unit Q: element k of this fibre is (a_0 = lane, a_1 = k, a_2); Q needs it when
                            // the column (lane, k) reaches beyond C
                            double *sc = scratch + size_t((it >> 27) & 31u) * a.ncolq + lane;
#pragma unroll
                            for (int k = 0; k < MAXT; ++k)
                                if (k < t2 && int(lane) < a.rq[k]) sc[a.cb[k]] = x[k];
                        }
                    }
                }
                stamp(sb + 4 + 4 * h);
                if (last_round && a.gather && have_next) issue_gather(nb, np);
            }
            if (!last) {
                __syncthreads();
                stamp(sb + 5 + 4 * h);
            }
        }
        if constexpr (TL) { if (p == 1) ++fn; }
    }
#undef S_BAR
#undef S_ITEM
#undef S_SEO
#undef S_PERM
#undef S_M
}

}  // namespace lpfnt
