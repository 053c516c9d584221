// Whole-function kernel (m = 2, 3; N < 65536): ONE pass over HBM per transform for batches of functions.
//
// Reference: itransform_lt_{2d,3d} / transform_lt_{2d,3d} / itransform_ut_* (src/lpfun/core/atoms.py:406,668,311,494,448,745),
// i.e. all m sweeps of a batched fnt / ifnt (src/lpfun/transform.py:461-475, 581-595) fused over one shared-memory copy.
//
//   * persistent CTAs: TG = 768 threads, one CTA per SM, for functions that fit only once into the 227 KB of an SM
//     (d=3, n=30: 119 KB), TG = 384 threads, two CTAs per SM, when two fit (the barrier / load / store phases of one CTA
//     then hide behind the arithmetic of the other);
//   * the function (N doubles, contiguous in HBM) arrives in shared memory by the bulk-copy engine (cp.async.bulk +
//     mbarrier, SASS UBLKCP), issued by one thread as soon as the CTA's previous function has left the buffer, i.e. it
//     overlaps the last sweep's arithmetic and stores;
//   * every sweep: thread <-> item (line for coordinate 0, element-fibre otherwise), fibre in registers, items in
//     length order (host), rounds of TG items; coordinates 0 .. m-2 in place in shared memory;
//   * the LAST coordinate goes from registers straight to HBM (fused colex scatter through a uint16 copy of the permutation
//     in shared memory); the colex gather of fnt is applied by the coordinate-0 loads the same way (single-round lists) or
//     by cp.async at prefetch time (multi-round lists, where an in-place gather would race).
// Algorithmic traffic = actual traffic = 16 B/DOF per transform.
//
// What round 2 measured and changed (profiles/r2_*):
//   * the unrolled triangle code alone reaches 93 % (exact) / 81 % (FMA) of the FP64 pipe with 24 warps per SM
//     (tools/probe4.cu); DMMA runs on the same pipe at 80 % of the DFMA rate, so tensor-core FP64 buys nothing here;
//   * ptxas hoists the ~500 loop-invariant constant loads of the triangle code out of the function / coordinate loops and
//     spills what does not fit into uniform registers (up to 2.7 GB of local-memory traffic per launch in one variant);
//     the matrix is therefore addressed through an offset the compiler cannot prove to be zero (W_M): 0 spills, which
//     also makes 4-row blocks (more independent chains per warp) fit into 80 registers;
//   * a variant that split every function into two units (planes a_2 < 13 / >= 13, partial folds handed over through an
//     L2-resident scratch buffer) to run two CTAs per SM at n = 30 was bit-exact but slower (1.56 vs 1.09 ms per fnt+ifnt):
//     no room left for the permutation table in shared memory, twice the barriers, 12-warp CTAs; removed
//     (profiles/r2_split_variant/).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "sweep_kernels.cuh"

namespace lpfnt {

constexpr int kWholeStamps = 16;     // clock64 stamps per warp and function (development timeline)

// item word: bits 0..15 = line offset (coordinate 0) or first entry of the fibre in seo (coordinate h >= 1),
//            bits 16..21 = length t (0 = empty slot), bits 22..31 = lane a_0 (coordinate h >= 1)
struct WholeArgs {
    const double *in;
    double *out;
    const uint32_t *items;             // all item lists back to back, in thread order
    const uint16_t *seo;               // per coordinate h >= 1 the line offsets in fibre order
    const uint16_t *perm16;            // colex permutation (or nullptr)
    long long *dbg;                    // development: clock64 stamps (nullptr in production)
    // PUSH instantiations: the result is stored into the SAME offset of every rank's copy of a symmetric buffer (an all-gather
    // done by the last sweep's stores, over NVLink): through the NVSwitch multicast mapping when there is one (one
    // multimem.st reaches every GPU, the own one included), else by one plain store per peer mapping
    double *mc_out;                    // multicast mapping of `out` (or nullptr)
    double *peer_out[8];               // peer mappings of `out`, the own one included (n_peers entries)
    int32_t n_peers;
    int32_t item_off[3], n_items[3];
    int32_t total_items, seo_len;
    int32_t gather, scatter;
    int32_t n_elem, m, batch;
};

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
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
__device__ __forceinline__ void multimem_st_f64(double *mc_addr, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc_addr), "d"(v) : "memory");
}
// one result element into every rank's copy; NOT inlined: the store code of k_whole is unrolled over every row of a fibre, and
// the per-peer loop at each of those sites made the PUSH instantiations 30 % of the library's code
static __device__ __noinline__ void whole_push_store(double *mc_out, double *const *peer_out, int n_peers, int64_t o, double v) {
    if (mc_out) multimem_st_f64(mc_out + o, v);
    else
        for (int r = 0; r < n_peers; ++r) peer_out[r][o] = v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr unsigned kBulkPiece = 32768;  // bytes per bulk-copy instruction

// shared memory of k_whole (bytes): data (+ 2 elements for the shifted origin of odd N), one mbarrier, items, seo, permutation
inline size_t whole_smem_bytes(int n_elem, int total_items, int seo_len, bool perm) {
    return size_t((n_elem + 3) & ~1) * 8 + 16 + size_t(total_items) * 4 + size_t((seo_len + 7) & ~7) * 2 + (perm ? size_t((n_elem + 7) & ~7) * 2 : 0);
}

// Two fibres per thread, sharing every matrix entry (lower mat-vec, rows k and k-1 of both fibres side by side: four
// independent chains per thread).  The constant path delivers one LDCU.128 (two entries) per ~8 cycles and SM
// sub-partition (tools/probe5.cu): enough for the exact form of ONE fibre (2 DMUL + 2 DADD = 8 pipe cycles per two
// entries), half of what the fused form needs (2 DFMA = 4 cycles) -- with two fibres per thread every entry feeds twice the
// arithmetic.  tg: warp maximum of both lengths.
template <int MAXT, bool FMA>
__device__ __forceinline__ void apply_fibre2(double (&x)[MAXT], double (&y)[MAXT], const PackedTri<MAXT> &M, const int tg) {
    static_assert(MAXT % 2 == 0, "rows are processed in pairs");
#pragma unroll
    for (int k = MAXT - 1; k >= 1; k -= 2) {
        if (k - 1 < tg) {
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
}

// NF = fibres per thread: 1 or 2 (the item lists then hold rounds of 2 * TG items, thread i owning slots i and TG + i).
// SEQ = false: the two fibres are folded side by side and share every matrix entry (lower mat-vec; two items of the same
// length class: the fused forms, whose constant path is the limit).  SEQ = true: one after the other, each with its own
// warp-uniform row bound -- the host pairs a LONG with a SHORT chunk of items per warp, so that all warps carry the same
// work (work is quadratic in the length: rank r + rank 23 - r is constant within 6 % at d = 3, n = 30), and the 168
// registers of a 384-thread CTA let ptxas keep the two row chains of the exact form interleaved (at the 80 registers of
// a 768-thread CTA it emits ONE serial DMUL -> DADD chain per warp: half the rate once a warp runs alone).
// MINB: CTAs per SM the register budget is cut for.
template <int MAXT, int MODE, bool FMA, int TG, int ROWS, int NF = 1, bool SEQ = false, int MINB = 1, bool TL = false, bool PUSH = false>
__global__ void __launch_bounds__(TG, MINB) k_whole(const __grid_constant__ WholeArgs a, const __grid_constant__ PackedTri<MAXT> M) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *const buf0 = reinterpret_cast<double *>(smem_raw);  // 16-byte aligned landing zone
#define W_NPAD ((a.n_elem + 3) & ~1)
#define W_BAR (reinterpret_cast<uint64_t *>(buf0 + W_NPAD))
#define W_SITEM (reinterpret_cast<uint32_t *>(buf0 + W_NPAD + 2))
#define W_SEO (reinterpret_cast<uint16_t *>(W_SITEM + a.total_items))
#define W_SPERM (W_SEO + ((a.seo_len + 7) & ~7))
    // The matrix parameter behind an offset the compiler cannot prove to be zero (r0 < 2^20 always): the triangle code sits
    // inside the function / coordinate / round loops, its constant loads are loop invariant, and ptxas hoists as many as
    // it finds uniform registers for and SPILLS the rest to local memory.  With the round counter in the address nothing is
    // invariant and nothing is hoisted.
#define W_M(r0) (*reinterpret_cast<const PackedTri<MAXT> *>(&M.v[((r0) >> 20) << 1]))
    const int N = a.n_elem;
    for (int i = threadIdx.x; i < a.total_items; i += TG) W_SITEM[i] = __ldg(a.items + i);
    for (int i = threadIdx.x; i < a.seo_len; i += TG) W_SEO[i] = __ldg(a.seo + i);
    if (a.gather || a.scatter)
        for (int e = threadIdx.x; e < N; e += TG) W_SPERM[e] = __ldg(a.perm16 + e);
    if (threadIdx.x == 0) {
        mbar_init(W_BAR, 1);
        mbar_init(W_BAR + 1, TG / 32);  // "every warp has gathered its lines" (one arrival per warp and function)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // colex gather of the input: by the coordinate-0 loads (every thread reads its line through the permutation, then a
    // barrier, then everybody writes) when coordinate 0 is ONE round; with several rounds that would race, and the
    // permutation is applied by cp.async while the function is fetched instead
    const bool gather_cp = a.gather && a.n_items[0] > TG * NF;
    const bool oddN = (N & 1) != 0;

    // Bulk copies move 16-byte granules between 16-byte aligned addresses.  For odd N every second function starts 8
    // bytes off: the copy then starts one element early (mis = 1) and the function sits at buf0 + mis; an even-b function
    // of an odd N copies one element more than it needs (the next function's first) -- except the very last function of
    // the batch, which copies N - 1 elements and fetches its last one with an ordinary load (nothing is read past the end).
    auto issue_bulk = [&](int b) {  // one thread
        const int mis = oddN ? (b & 1) : 0;
        const bool tail = oddN && mis == 0 && b == a.batch - 1;
        const unsigned nbytes = unsigned(oddN ? (tail ? N - 1 : N + 1) : N) * 8u;
        const char *src = reinterpret_cast<const char *>(a.in + int64_t(b) * N - mis);
        mbar_expect_tx(W_BAR, nbytes);
        for (unsigned o = 0; o < nbytes; o += kBulkPiece)
            bulk_g2s(reinterpret_cast<char *>(buf0) + o, src + o, min(kBulkPiece, nbytes - o), W_BAR);
        if (tail) buf0[N - 1] = a.in[int64_t(b) * N + N - 1];  // ordered before the consumers by the barrier after the mbarrier wait
    };
    auto issue_gather = [&](int b) {  // every thread its share
        const double *src = a.in + int64_t(b) * N;
#pragma unroll 4
        for (int i = threadIdx.x; i < N; i += TG) cp_async8(buf0 + i, src + W_SPERM[i]);
        cp_async_commit();
    };

    int fn = 0;  // TL: development timeline (separate instantiation; the production kernel carries none of this)
    auto stamp = [&](int id) {
        if constexpr (TL) {
            if (a.dbg && blockIdx.x < 2 && fn < 4 && (threadIdx.x & 31) == 0) {
                long long c;
                asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
                a.dbg[((fn * 2 + blockIdx.x) * (TG / 32) + (threadIdx.x >> 5)) * kWholeStamps + id] = c;
            }
        }
    };

    if (int(blockIdx.x) < a.batch) {
        if (gather_cp) issue_gather(blockIdx.x);
        else if (threadIdx.x == 0) issue_bulk(blockIdx.x);
    }
    unsigned parity = 0, gparity = 0;
#pragma unroll 1
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        stamp(0);
        if (gather_cp) {
            cp_async_wait_all();
        } else {
            mbar_wait(W_BAR, parity);
            parity ^= 1;
        }
        if (gather_cp || (oddN && b == a.batch - 1)) __syncthreads();  // cp.async data / the last element stored by thread 0
        stamp(1);
#pragma unroll 1
        for (int h = 0; h < a.m; ++h) {
            const bool last = (h == a.m - 1);
#pragma unroll 1
            for (int r0 = 0; r0 < a.n_items[h]; r0 += TG * NF) {
                static_assert(NF == 1 || (NF == 2 && (SEQ || MODE == kLowerMatvec)), "two fibres side by side: lower mat-vec only");
                double x[NF][MAXT];
                int tmax = 0, tf[NF];
                {
                    double *buf = buf0 + ((oddN && !gather_cp) ? (b & 1) : 0);
                    uint32_t it[NF];
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        it[f] = W_SITEM[a.item_off[h] + r0 + f * TG + threadIdx.x];
                        tf[f] = int((it[f] >> 16) & 63u);
                        tmax = max(tmax, tf[f]);
                    }
                    if (h == 0 && a.gather && !gather_cp) {
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            const int t = int((it[f] >> 16) & 63u);
                            const uint16_t *sp = W_SPERM + (it[f] & 0xffffu);
#pragma unroll
                            for (int k = 0; k < MAXT; ++k) x[f][k] = (k < t) ? buf[sp[k]] : 0.0;
                        }
                        // The gather reads other lines' slots: nobody may write before everybody has read.  As a SPLIT barrier
                        // (arrive here, wait in front of the stores): a CTA barrier at this place held the warps that finish
                        // the previous function's last sweep early -- those with the short items -- until the last warp had
                        // arrived (9 K of 42.6 K cycles per function in the phase timeline).
                        __syncwarp();
                        if ((threadIdx.x & 31) == 0) mbar_arrive(W_BAR + 1);
                    } else if (h == 0) {
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            const int t = int((it[f] >> 16) & 63u);
                            const double *q = buf + (it[f] & 0xffffu);
#pragma unroll
                            for (int k = 0; k < MAXT; ++k) x[f][k] = (k < t) ? q[k] : 0.0;
                        }
                    } else {
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            const int t = int((it[f] >> 16) & 63u);
                            // the fibre's row offsets are 8-byte aligned in seo (host): four per shared-memory load
                            const uint2 *eo4 = reinterpret_cast<const uint2 *>(W_SEO + (it[f] & 0xffffu));
                            const double *q = buf + (it[f] >> 22);
#pragma unroll
                            for (int k4 = 0; k4 < MAXT; k4 += 4) {
                                const uint2 e = eo4[k4 >> 2];
                                x[f][k4] = (k4 < t) ? q[e.x & 0xffffu] : 0.0;
                                x[f][k4 + 1] = (k4 + 1 < t) ? q[e.x >> 16] : 0.0;
                                x[f][k4 + 2] = (k4 + 2 < t) ? q[e.y & 0xffffu] : 0.0;
                                x[f][k4 + 3] = (k4 + 3 < t) ? q[e.y >> 16] : 0.0;
                            }
                        }
                    }
                }
                if (last && r0 + TG * NF >= a.n_items[h]) {
                    // every thread holds its fibres in registers: the buffer is free for the CTA's next function
                    fence_proxy_async();
                    __syncthreads();
                    if (!gather_cp && threadIdx.x == 0 && b + int(gridDim.x) < a.batch) issue_bulk(b + gridDim.x);
                }
                stamp(2 + 4 * h);
                if constexpr (NF == 2 && !SEQ) {
                    apply_fibre2<MAXT, FMA>(x[0], x[1], W_M(r0), __reduce_max_sync(0xffffffffu, tmax));
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) apply_fibre<MAXT, MODE, FMA, ROWS>(x[f], tf[f], W_M(r0 + f), __reduce_max_sync(0xffffffffu, tf[f]));
                }
                stamp(3 + 4 * h);
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    // the item is read again: nothing but the fibres stays in registers across the triangle code
                    // (in particular not the "k < t" predicates, which the compiler would pack into a bit mask)
                    const uint32_t it = *reinterpret_cast<const volatile uint32_t *>(W_SITEM + a.item_off[h] + r0 + f * TG + threadIdx.x);
                    const int t2 = int((it >> 16) & 63u);
                    if (last) {
                        const uint2 *eo4 = reinterpret_cast<const uint2 *>(W_SEO + (it & 0xffffu));
                        const unsigned add = it >> 22;
                        double *out = a.out + int64_t(b) * N;
                        // one result element -> out[idx] (PUSH: the same offset on every rank)
                        auto put = [&](unsigned idx, double v) {
                            if constexpr (PUSH) {
                                const int64_t o = int64_t(b) * N + idx;
                                whole_push_store(a.mc_out, a.peer_out, a.n_peers, o, v);
                            } else {
                                out[idx] = v;
                            }
                        };
                        if (a.scatter) {
                            // blocks of eight rows: offsets (two loads), permutation look-ups (eight independent loads),
                            // stores -- one dependent chain per block instead of one per element
#pragma unroll
                            for (int k8 = 0; k8 < MAXT; k8 += 8) {
                                const uint2 e0 = eo4[k8 >> 2], e1 = eo4[(k8 >> 2) + 1];
                                const unsigned o[8] = {e0.x & 0xffffu, e0.x >> 16, e0.y & 0xffffu, e0.y >> 16,
                                                       e1.x & 0xffffu, e1.x >> 16, e1.y & 0xffffu, e1.y >> 16};
                                unsigned u[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) u[j] = (k8 + j < t2) ? W_SPERM[o[j] + add] : 0u;  // rows past the fibre hold padding
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (k8 + j < t2) put(u[j], x[f][k8 + j]);
                            }
                        } else {
#pragma unroll
                            for (int k4 = 0; k4 < MAXT; k4 += 4) {
                                const uint2 e = eo4[k4 >> 2];
                                if (k4 < t2) put((e.x & 0xffffu) + add, x[f][k4]);
                                if (k4 + 1 < t2) put((e.x >> 16) + add, x[f][k4 + 1]);
                                if (k4 + 2 < t2) put((e.y & 0xffffu) + add, x[f][k4 + 2]);
                                if (k4 + 3 < t2) put((e.y >> 16) + add, x[f][k4 + 3]);
                            }
                        }
                    } else {
                        double *buf = buf0 + ((oddN && !gather_cp) ? (b & 1) : 0);
                        if (h == 0) {
                            if (f == 0 && a.gather && !gather_cp) {  // every warp has read its lines through the permutation
                                mbar_wait(W_BAR + 1, gparity);
                                gparity ^= 1;
                            }
                            double *q = buf + (it & 0xffffu);
#pragma unroll
                            for (int k = 0; k < MAXT; ++k)
                                if (k < t2) q[k] = x[f][k];
                        } else {
                            const uint2 *eo4 = reinterpret_cast<const uint2 *>(W_SEO + (it & 0xffffu));
                            double *q = buf + (it >> 22);
#pragma unroll
                            for (int k4 = 0; k4 < MAXT; k4 += 4) {
                                const uint2 e = eo4[k4 >> 2];
                                if (k4 < t2) q[e.x & 0xffffu] = x[f][k4];
                                if (k4 + 1 < t2) q[e.x >> 16] = x[f][k4 + 1];
                                if (k4 + 2 < t2) q[e.y & 0xffffu] = x[f][k4 + 2];
                                if (k4 + 3 < t2) q[e.y >> 16] = x[f][k4 + 3];
                            }
                        }
                    }
                }
                stamp(4 + 4 * h);
                // the permuting prefetch is issued by every thread once its own fibre is dead (issued in front of the
                // triangle code it would push the 64-register fibre into local memory)
                if (last && gather_cp && r0 + TG * NF >= a.n_items[h] && b + int(gridDim.x) < a.batch) issue_gather(b + gridDim.x);
            }
            if (!last) {
                __syncthreads();
                stamp(5 + 4 * h);
            }
        }
        if constexpr (TL) ++fn;
    }
#undef W_NPAD
#undef W_BAR
#undef W_SITEM
#undef W_SEO
#undef W_SPERM
#undef W_M
}

}  // namespace lpfnt
