// Point evaluation of a Newton-basis expansion on a lower set (sm_100a).
//
// Reference: newton2point, src/lpfun/utils.py:132-162 -- per point a basis table, then
// sum_r c[r] * prod_j basis[j][A[r,j]]  (N*(m+1) flops per point, int64 gathers of A).
// Here the sum is factorised over the index trie and evaluated in nested Horner form:
//   S_1(line)  = sum_{a0} c[line,a0] prod_{i<a0}(x_0 - node_i)      Horner over a0
//   S_{j+1}    = sum_{a_j} S_j(child a_j) prod_{i<a_j}(x_j - node_i) Horner over a_j
// One FMA per coefficient plus one per trie node (~1.09 N per point at d=4,n=20).  The
// operation order differs from the reference's flat sum; eval is well conditioned (SURVEY.md
// 0.3: <= 2e-14 relative) and is checked against the oracle with a stated tolerance.
//
// Thread <-> point(s).  All threads walk the lines in reverse storage order in lock step, so
// every coefficient load is warp-uniform (broadcast) and control flow never diverges.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lpfnt {

template <int MAXT>
struct NodeTable {
    double v[MAXT];
};

struct EvalArgs {
    const double *coeffs;      // N
    const double *points;      // P x m
    double *values;            // P
    const int32_t *line_off;   // N_1 + 1
    const uint8_t *line_alpha; // N_1 x (m-1)
    int64_t num_points;
    int32_t num_lines;
    int32_t m;
};

constexpr int kEvalMaxM = 12;

template <int MAXT, int PPT, int MM, int THREADS>
__global__ void __launch_bounds__(THREADS) k_eval_horner(const EvalArgs a, const __grid_constant__ NodeTable<MAXT> nodes) {
    const int64_t p0 = (int64_t(blockIdx.x) * THREADS + threadIdx.x) * PPT;
    const int m = a.m;
    double x[PPT][MM];
    double d0[PPT][MAXT];
    double acc[PPT][MM];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int64_t p = min(p0 + q, a.num_points - 1);
#pragma unroll
        for (int j = 0; j < MM; ++j) {
            x[q][j] = (j < m) ? a.points[p * m + j] : 0.0;
            acc[q][j] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < MAXT; ++k) d0[q][k] = x[q][0] - nodes.v[k];
    }
    for (int l = a.num_lines - 1; l >= 0; --l) {
        const int off = a.line_off[l];
        const int t = a.line_off[l + 1] - off;
        const double *__restrict__ c = a.coeffs + off;
        // every coefficient of the line is requested before the first FMA needs one (the loads are
        // warp-uniform broadcasts); the Horner chain then runs in uniform blocks of four
        double cv[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) cv[k] = (k < t) ? __ldg(c + k) : 0.0;
        double s[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) s[q] = 0.0;
#pragma unroll
        for (int kb = MAXT - 4; kb >= 0; kb -= 4) {
            if (kb < t) {  // zero padding: s stays 0 until the first real coefficient
#pragma unroll
                for (int k = kb + 3; k >= kb; --k) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) s[q] = __fma_rn(s[q], d0[q][k], cv[k]);
                }
            }
        }
        // fold the finished line into level 1, and every level that completes with it upward
        const uint8_t *al = a.line_alpha + int64_t(l) * (m - 1);
        bool carry = true;
#pragma unroll
        for (int j = 1; j < MM; ++j) {
            if (j < m && carry) {
                const int aj = al[j - 1];
                const double nd = nodes.v[aj];
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    const double below = (j == 1) ? s[q] : acc[q][j - 1];
                    acc[q][j] = __fma_rn(acc[q][j], x[q][j] - nd, below);
                    if (j > 1) acc[q][j - 1] = 0.0;
                }
                carry = (aj == 0);
            }
        }
        if (m == 1) {
#pragma unroll
            for (int q = 0; q < PPT; ++q) acc[q][0] = s[q];
        }
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        double v = acc[q][0];
#pragma unroll
        for (int j = 1; j < MM; ++j)
            if (j == m - 1) v = acc[q][j];
        if (p0 + q < a.num_points) a.values[p0 + q] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Block-staged evaluation (the fast path): shared staging of k_eval_tree below.
//   * the coefficients are first re-laid out with every line padded to a multiple of 4 doubles
//     (k_eval_pad; zero filled), so a line is a run of aligned 32-byte groups;
//   * a CTA stages blocks of <= 256 lines (<= 24 KB of padded coefficients + descriptors) in shared
//     memory with cp.async, double buffered: no global latency on the critical path;
//   * coefficients come from shared memory as 16-byte broadcasts (one LDS per two coefficients).
// (Round 1's kernel on this staging, k_eval_block -- two lines at a time, a length test per four coefficients -- was replaced
// by k_eval_tree in round 2 and removed: 122 -> 91 ms Newton, 202 -> 113 ms Chebyshev at d = 4, n = 20, 10^7 points.)
// ---------------------------------------------------------------------------------------
struct EvalBlockArgs {
    const double *cpad;        // padded coefficients
    const double *points;      // P x m
    double *values;            // P
    const int4 *blocks;        // {first line, #lines, padded offset, padded size}
    const int2 *lines;         // {padded offset, t}
    const int2 *tlines;        // k_eval_tree: {byte offset in the staged block, steps | 8 a_1 << 8 | depth << 16}
    const uint8_t *line_alpha; // N_1 x (m-1)
    int64_t num_points;
    int32_t num_blocks;
    int32_t m;
};

constexpr int kEvalLB = 256;    // host: kEvalBlockLines
constexpr int kEvalCB = 2048;   // host: kEvalBlockElems

struct EvalPadArgs {
    const double *coeffs;
    double *cpad;
    const int32_t *line_off;
    const int2 *lines;
    int32_t num_lines;
};

#ifdef LPFNT_EVAL_TU
__global__ void __launch_bounds__(128) k_eval_pad(const EvalPadArgs a) {
    const int l = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (l >= a.num_lines) return;
    const int off = a.line_off[l], t = a.line_off[l + 1] - off;
    const int2 ln = a.lines[l];
    const int tp = (t + 3) & ~3;
    for (int k = lane; k < tp; k += 32) a.cpad[ln.x + k] = (k < t) ? a.coeffs[off + k] : 0.0;
}
#endif

__device__ __forceinline__ void cp_async16_e(void *smem_dst, const void *gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8_e(void *smem_dst, const void *gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async1_e(void *smem_dst, const void *gsrc, int bytes4) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc));
    (void)bytes4;
}

// ---------------------------------------------------------------------------------------
// Evaluation, round 2: the nested Horner over the index trie rebuilt around the instruction count.  ncu of round 1's kernel
// (profiles/r1_k_eval_block_*): 72 % of the issued instructions were not the coefficient FMAs -- per line ~56 instructions of
// length tests (one compare + branch per four coefficients and line), descriptor / alpha look-ups and a fold loop over all m
// levels with carry flags -- against 2 x 11 FMAs, at 45 % issue utilisation on three warps per sub-partition.  Here
//   * ONE line at a time, entered through a jump table on ceil(t / 4) into a chain that falls through: no per-step tests;
//   * the descriptor carries a_1 and the fold depth: level 1 is folded with one indexed constant load, and only a line with
//     a_1 = 0 (one per group of lines) walks the upper levels;
//   * the next step's coefficients are fetched before the current step's FMAs (two register sets used in turn).
// Every line's chain is the one round 1 ran, so the values are unchanged bit for bit.
// CHEB = true evaluates in the Chebyshev basis (chebyshev2point, src/lpfun/utils.py:165-195) with the same trie walk:
//   * coordinate 0: d0[q][k] holds T_k(x_0) (three-term recurrence, as the reference builds its basis table) and a
//     line is the dot product sum_k c_k T_k(x_0) (the zero padding contributes nothing);
//   * coordinates j >= 1: the lines are visited with a_j DEscending, so sum_{a_j} T_{a_j}(x_j) * below_{a_j} is a
//     Clenshaw recurrence b = below + 2 x_j b1 - b2 with two running values per level; at a_j = 0 the level's value
//     is b - x_j b1 and the state resets.
// ---------------------------------------------------------------------------------------
template <int MAXT, int PPT, int MM, int THREADS, int MINB, bool CHEB = false>
__global__ void __launch_bounds__(THREADS, MINB) k_eval_tree(const EvalBlockArgs a, const __grid_constant__ NodeTable<MAXT> nodes) {
    __shared__ __align__(16) double scb[2][kEvalCB];
    __shared__ int2 sln[2][kEvalLB];
    __shared__ __align__(4) uint8_t sal[2][kEvalLB * (MM - 1) + 4];
    const int tid = threadIdx.x;
    const int m = a.m;
    const int64_t p0 = (int64_t(blockIdx.x) * THREADS + tid) * PPT;
    double x[PPT][MM], d0[PPT][MAXT], acc[PPT][MM];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int64_t p = min(p0 + q, a.num_points - 1);
#pragma unroll
        for (int j = 0; j < MM; ++j) {
            x[q][j] = (j < m) ? a.points[p * m + j] : 0.0;
            acc[q][j] = 0.0;
        }
        if constexpr (CHEB) {  // T_k(x_0) by the three-term recurrence, as the reference builds its basis table (utils.py:165-195)
            d0[q][0] = 1.0;
            if (MAXT > 1) d0[q][1] = x[q][0];
#pragma unroll
            for (int k = 1; k + 1 < MAXT; ++k) d0[q][k + 1] = __dsub_rn(__dmul_rn(__dmul_rn(2.0, x[q][0]), d0[q][k]), d0[q][k - 1]);
        } else {
#pragma unroll
            for (int k = 0; k < MAXT; ++k) d0[q][k] = x[q][0] - nodes.v[k];
        }
    }
    // CHEB: levels j >= 1 are Clenshaw recurrences (the lines come with a_j descending): acc[][] holds b_{k+1}, b2[][] b_{k+2}
    double b2[CHEB ? PPT : 1][CHEB ? MM : 1], res[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) res[q] = 0.0;
    if constexpr (CHEB) {
#pragma unroll
        for (int q = 0; q < PPT; ++q)
#pragma unroll
            for (int j = 0; j < MM; ++j) b2[q][j] = 0.0;
    }

    auto issue = [&](int bi) {  // stage block bi into buffer bi & 1
        const int4 b = a.blocks[bi];
        const int s = bi & 1;
        for (int i = tid; 2 * i < b.w; i += THREADS) cp_async16_e(&scb[s][2 * i], a.cpad + b.z + 2 * i);
        for (int i = tid; i < b.y; i += THREADS) cp_async8_e(&sln[s][i], a.tlines + b.x + i);
        if (m > 2) {  // alpha bytes (only the lines with a_1 = 0 read them): 4-byte words from an aligned start
            const int64_t a0 = int64_t(b.x) * (m - 1);
            const int shift = int(a0 & 3);
            const int nbytes = b.y * (m - 1) + shift;
            const uint8_t *src = a.line_alpha + (a0 - shift);
            for (int i = tid; 4 * i < nbytes; i += THREADS) cp_async1_e(&sal[s][4 * i], src + 4 * i, 0);
        }
        asm volatile("cp.async.commit_group;" ::);
    };

    issue(a.num_blocks - 1);
    for (int bi = a.num_blocks - 1; bi >= 0; --bi) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // block bi landed; everybody is done with the buffer block bi-1 goes into
        if (bi > 0) issue(bi - 1);
        const int4 b = a.blocks[bi];
        const int s = bi & 1;
        const int shift = int((int64_t(b.x) * (m - 1)) & 3);
        const uint8_t *al_base = &sal[s][shift];
        const char *cbase = reinterpret_cast<const char *>(&scb[s][0]);
        int2 ln = sln[s][b.y - 1];
        for (int li = b.y - 1; li >= 0; --li) {
            // descriptor: x = byte offset of the line in the staged block; y = steps | 8 a_1 << 8 | depth << 16
            const int steps = ln.y & 255, a1x8 = (ln.y >> 8) & 255, depth = ln.y >> 16;
            const double *c = reinterpret_cast<const double *>(cbase + ln.x);
            if (li > 0) ln = sln[s][li - 1];  // next line's descriptor: in flight during this line
            double sl[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) sl[q] = 0.0;
            // two coefficient register sets used in turn (even / odd steps): the next step's coefficients are fetched before
            // the current step's FMAs without register moves; every entry of the jump table fetches its own first step
            double2 A01, A23, B01, B23;
#define LPFNT_EVAL_LOAD(KB, C01, C23)                                                                \
    if constexpr ((KB) < MAXT) {                                                                     \
        C01 = *reinterpret_cast<const double2 *>(c + (KB));                                          \
        C23 = *reinterpret_cast<const double2 *>(c + (KB) + 2);                                      \
    }
#define LPFNT_EVAL_STEP(KB, C01, C23, N01, N23)                                                      \
    if constexpr ((KB) < MAXT) {                                                                     \
        if constexpr ((KB) > 0) {                                                                    \
            N01 = *reinterpret_cast<const double2 *>(c + (KB) - 4);                                  \
            N23 = *reinterpret_cast<const double2 *>(c + (KB) - 2);                                  \
        }                                                                                            \
        _Pragma("unroll") for (int q = 0; q < PPT; ++q) {                                            \
            if constexpr (CHEB) {                                                                    \
                sl[q] = __fma_rn(C23.y, d0[q][(KB) + 3 < MAXT ? (KB) + 3 : 0], sl[q]);               \
                sl[q] = __fma_rn(C23.x, d0[q][(KB) + 2 < MAXT ? (KB) + 2 : 0], sl[q]);               \
                sl[q] = __fma_rn(C01.y, d0[q][(KB) + 1 < MAXT ? (KB) + 1 : 0], sl[q]);               \
                sl[q] = __fma_rn(C01.x, d0[q][(KB) < MAXT ? (KB) : 0], sl[q]);                       \
            } else {                                                                                 \
                sl[q] = __fma_rn(sl[q], d0[q][(KB) + 3 < MAXT ? (KB) + 3 : 0], C23.y);               \
                sl[q] = __fma_rn(sl[q], d0[q][(KB) + 2 < MAXT ? (KB) + 2 : 0], C23.x);               \
                sl[q] = __fma_rn(sl[q], d0[q][(KB) + 1 < MAXT ? (KB) + 1 : 0], C01.y);               \
                sl[q] = __fma_rn(sl[q], d0[q][(KB) < MAXT ? (KB) : 0], C01.x);                       \
            }                                                                                        \
        }                                                                                            \
    }
            switch (steps) {
                case 8: LPFNT_EVAL_LOAD(28, B01, B23) goto step7;
                case 7: LPFNT_EVAL_LOAD(24, A01, A23) goto step6;
                case 6: LPFNT_EVAL_LOAD(20, B01, B23) goto step5;
                case 5: LPFNT_EVAL_LOAD(16, A01, A23) goto step4;
                case 4: LPFNT_EVAL_LOAD(12, B01, B23) goto step3;
                case 3: LPFNT_EVAL_LOAD(8, A01, A23) goto step2;
                case 2: LPFNT_EVAL_LOAD(4, B01, B23) goto step1;
                default: LPFNT_EVAL_LOAD(0, A01, A23) goto step0;
            }
        step7: LPFNT_EVAL_STEP(28, B01, B23, A01, A23)
        step6: LPFNT_EVAL_STEP(24, A01, A23, B01, B23)
        step5: LPFNT_EVAL_STEP(20, B01, B23, A01, A23)
        step4: LPFNT_EVAL_STEP(16, A01, A23, B01, B23)
        step3: LPFNT_EVAL_STEP(12, B01, B23, A01, A23)
        step2: LPFNT_EVAL_STEP(8, A01, A23, B01, B23)
        step1: LPFNT_EVAL_STEP(4, B01, B23, A01, A23)
        step0: LPFNT_EVAL_STEP(0, A01, A23, B01, B23)
#undef LPFNT_EVAL_STEP
#undef LPFNT_EVAL_LOAD
            if constexpr (MM == 1) {  // m = 1 (dispatched to this instantiation): the single line is the value
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[q][0] = sl[q];
            } else if constexpr (CHEB) {
                // Clenshaw step of level 1 with the line's value; a_1 = 0 (depth > 0) completes the level: its value goes up
                double bq[PPT];
#pragma unroll
                for (int q = 0; q < PPT; ++q) bq[q] = __fma_rn(2.0 * x[q][1], acc[q][1], sl[q] - b2[q][1]);
                if (depth == 0) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) { b2[q][1] = acc[q][1]; acc[q][1] = bq[q]; }
                } else {
                    double val[PPT];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        val[q] = __fma_rn(-x[q][1], acc[q][1], bq[q]);
                        if (m == 2) res[q] = val[q];
                        acc[q][1] = 0.0;
                        b2[q][1] = 0.0;
                    }
                    const uint8_t *al = al_base + li * (m - 1);
                    bool carry = true;
#pragma unroll
                    for (int j = 2; j < MM; ++j) {
                        if (j < m && carry) {
                            const int aj = al[j - 1];
#pragma unroll
                            for (int q = 0; q < PPT; ++q) {
                                const double b = __fma_rn(2.0 * x[q][j], acc[q][j], val[q] - b2[q][j]);
                                if (aj == 0) {
                                    val[q] = __fma_rn(-x[q][j], acc[q][j], b);
                                    if (j == m - 1) res[q] = val[q];  // the root: once, with the very last line
                                    acc[q][j] = 0.0;
                                    b2[q][j] = 0.0;
                                } else {
                                    b2[q][j] = acc[q][j];
                                    acc[q][j] = b;
                                }
                            }
                            carry = (aj == 0);
                        }
                    }
                }
            } else {
                const double nd1 = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(nodes.v) + a1x8);
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[q][1] = __fma_rn(acc[q][1], x[q][1] - nd1, sl[q]);
                if (depth > 0) {  // a_1 = 0: level 1 is complete and goes up, with every further level that completes
                    const uint8_t *al = al_base + li * (m - 1);
                    bool carry = true;
#pragma unroll
                    for (int j = 2; j < MM; ++j) {
                        if (j < m && carry) {
                            const int aj = al[j - 1];
                            const double nd = nodes.v[aj];
#pragma unroll
                            for (int q = 0; q < PPT; ++q) {
                                acc[q][j] = __fma_rn(acc[q][j], x[q][j] - nd, acc[q][j - 1]);
                                acc[q][j - 1] = 0.0;
                            }
                            carry = (aj == 0);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        double v = acc[q][0];
#pragma unroll
        for (int j = 1; j < MM; ++j)
            if (j == m - 1) v = acc[q][j];
        if constexpr (CHEB && MM > 1) v = res[q];
        if (p0 + q < a.num_points) a.values[p0 + q] = v;
    }
}


// ---------------------------------------------------------------------------------------
// Few points (P << 256 * #SM): the kernels above give every point ONE thread that walks all N coefficients -- eval of the
// README's "10 random points" at d = 6, n = 20 ran 240 ms on one CTA.  Split path: the index trie is cut below its top L
// levels into G subtrees (fixed (a_{m-L}, .., a_{m-1}); contiguous runs of lines, since the lines are in colex order);
//   k_eval_sub     : thread <-> (subtree, point): nested Horner over the subtree's lines -> partial[g][p]
//   k_eval_combine : thread <-> point: value = sum_g partial[g][p] * prod_{j >= m-L} N_{a_j(g)}(x_j)   (basis form, in g order)
// Deterministic (no atomics).  Newton basis.
// ---------------------------------------------------------------------------------------
struct EvalSubArgs {
    const double *coeffs;       // N
    const double *points;       // P x m
    double *partial;            // G x P
    const int32_t *line_off;    // N_1 + 1
    const uint8_t *line_alpha;  // N_1 x (m-1)
    const int32_t *sub_l0;      // G + 1: first line of every subtree
    const uint8_t *sub_alpha;   // G x L: (a_{m-L}, .., a_{m-1}) of every subtree
    double *values;             // P
    int64_t num_points;
    int32_t num_sub, L, m, n1;
};

#ifdef LPFNT_EVAL_TU
template <int MAXT>
__global__ void __launch_bounds__(128) k_eval_sub(const EvalSubArgs a, const __grid_constant__ NodeTable<MAXT> nodes) {
    const int64_t idx = int64_t(blockIdx.x) * 128 + threadIdx.x;
    if (idx >= int64_t(a.num_sub) * a.num_points) return;
    const int g = int(idx / a.num_points);
    const int64_t p = idx - int64_t(g) * a.num_points;
    const int m = a.m, lev = m - a.L;  // levels 0 .. lev-1 are folded here
    double x[kEvalMaxM], acc[kEvalMaxM];
    for (int j = 0; j < lev; ++j) { x[j] = a.points[p * m + j]; acc[j] = 0.0; }
    double d0[MAXT];
#pragma unroll
    for (int k = 0; k < MAXT; ++k) d0[k] = x[0] - nodes.v[k];
    for (int l = a.sub_l0[g + 1] - 1; l >= a.sub_l0[g]; --l) {
        const int off = a.line_off[l], t = a.line_off[l + 1] - off;
        // all coefficients of the line are requested before the Horner chain needs the first one
        double cv[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; ++k) cv[k] = (k < t) ? __ldg(a.coeffs + off + k) : 0.0;
        double s = 0.0;
#pragma unroll
        for (int k = MAXT - 1; k >= 0; --k)
            if (k < t) s = __fma_rn(s, d0[k], cv[k]);
        const uint8_t *al = a.line_alpha + int64_t(l) * (m - 1);
        bool carry = true;
        for (int j = 1; j < lev && carry; ++j) {
            const int aj = al[j - 1];
            const double below = (j == 1) ? s : acc[j - 1];
            acc[j] = __fma_rn(acc[j], x[j] - nodes.v[aj], below);
            if (j > 1) acc[j - 1] = 0.0;
            carry = (aj == 0);
        }
        if (lev == 1) acc[0] = s;  // one line per subtree
    }
    a.partial[idx] = acc[lev - 1];
}

// one CTA of 256 threads per point: thread i sums the subtrees i, i + 256, ... in order, then a fixed-shape tree over the
// 256 partial sums (deterministic)
template <int MAXT>
__global__ void __launch_bounds__(256) k_eval_combine(const EvalSubArgs a, const __grid_constant__ NodeTable<MAXT> nodes) {
    __shared__ double B[kEvalMaxM][MAXT];  // Newton basis values N_k(x_j) of the top L coordinates (utils.py:148-151)
    __shared__ double red[256];
    const int64_t p = blockIdx.x;
    const int m = a.m, L = a.L;
    if (threadIdx.x < L) {
        const int j = threadIdx.x;
        const double xj = a.points[p * m + (m - L + j)];
        B[j][0] = 1.0;
        for (int k = 1; k < a.n1; ++k) B[j][k] = B[j][k - 1] * (xj - nodes.v[k - 1]);
    }
    __syncthreads();
    double v = 0.0;
    for (int g = threadIdx.x; g < a.num_sub; g += 256) {
        double w = a.partial[int64_t(g) * a.num_points + p];
        for (int j = 0; j < L; ++j) w *= B[j][a.sub_alpha[g * L + j]];
        v += w;
    }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.values[p] = red[0];
}
#endif

// Fallback without unrolling limits (any n, any m): flat sum in the reference's own order,
// thread per point, basis table in local memory.  A: int64 (N x m) multi-indices on the device.
struct EvalGenericArgs {
    const double *coeffs, *points, *nodes;
    double *values;
    const int64_t *A;
    double *scratch;  // num_points * m * (n+1) doubles
    int64_t num_points, N;
    int32_t m, n;
    int32_t chebyshev;  // 1: T_k basis by the three-term recurrence (chebyshev2point, utils.py:165-195)
};

#ifdef LPFNT_EVAL_TU  // defined once, in eval_inst.cu
__global__ void __launch_bounds__(128) k_eval_generic(const EvalGenericArgs a) {
    const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= a.num_points) return;
    const int m = a.m, n1 = a.n + 1;
    double *basis = a.scratch + p * int64_t(m) * n1;
    for (int i = 0; i < m; ++i) {
        const double xi = a.points[p * m + i];
        basis[i * n1] = 1.0;
        if (a.chebyshev) {
            if (n1 > 1) basis[i * n1 + 1] = xi;
            for (int q = 1; q + 1 < n1; ++q)
                basis[i * n1 + q + 1] = __dsub_rn(__dmul_rn(__dmul_rn(2.0, xi), basis[i * n1 + q]), basis[i * n1 + q - 1]);
        } else {
            for (int q = 1; q < n1; ++q) basis[i * n1 + q] = __dmul_rn(basis[i * n1 + q - 1], __dsub_rn(xi, a.nodes[q - 1]));
        }
    }
    double value = 0.0;
    for (int64_t r = 0; r < a.N; ++r) {
        double prod = 1.0;
        for (int jj = 0; jj < m; ++jj) prod = __dmul_rn(prod, basis[jj * n1 + a.A[r * m + jj]]);
        value = __dadd_rn(value, __dmul_rn(a.coeffs[r], prod));
    }
    a.values[p] = value;
}

// FP64 FMA throughput probe: 8 independent dependent-chains per thread.
__global__ void __launch_bounds__(256) k_fp64_probe(double *sink, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, b, c); a1 = __fma_rn(a1, b, c); a2 = __fma_rn(a2, b, c); a3 = __fma_rn(a3, b, c);
        a4 = __fma_rn(a4, b, c); a5 = __fma_rn(a5, b, c); a6 = __fma_rn(a6, b, c); a7 = __fma_rn(a7, b, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) sink[0] = r;
}

#endif

}  // namespace lpfnt
