// Instantiates the register-fibre sweep kernels for one fibre capacity LPFNT_MAXT.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sweep_launch.h"

#ifndef LPFNT_MAXT
#error "compile with -DLPFNT_MAXT=<8|16|24|32>"
#endif

namespace lpfnt {

namespace {
constexpr int T = LPFNT_MAXT;

template <int MODE, bool FMA>
cudaError_t line_launch(const LineSweepArgs &a, const PackedTri<T> &M, int batch, cudaStream_t s) {
    dim3 grid((a.num_lines + kSweepThreads - 1) / kSweepThreads, batch);
    k_sweep_lines<T, MODE, FMA, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    return cudaGetLastError();
}

template <int MODE, bool FMA>
cudaError_t group_launch(const GroupSweepArgs &a, const PackedTri<T> &M, int batch, cudaStream_t s) {
    dim3 grid((kRunLanes * int64_t(a.num_groups) + kSweepThreads - 1) / kSweepThreads, batch);
    k_sweep_groups<T, MODE, FMA, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    return cudaGetLastError();
}

template <int MODE, bool FMA, int TG, int ROWS, int NF = 1, bool SEQ = false, int MINB = 1, bool TL = false, bool PUSH = false>
cudaError_t whole_launch(const WholeArgs &a, const PackedTri<T> &M, int grid, size_t smem, bool set_attr, cudaStream_t s) {
    auto kern = k_whole<T, MODE, FMA, TG, ROWS, NF, SEQ, MINB, TL, PUSH>;
    if (set_attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, TG, smem, s>>>(a, M);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && std::getenv("LPFNT_DEBUG_OCCUPANCY")) std::fprintf(stderr, "k_whole<%d, mode %d, fma %d, %d thr, NF %d, push %d> launch: %s (grid %d, smem %zu)\n", T, MODE, int(FMA), TG, NF, int(PUSH), cudaGetErrorString(e), grid, smem);
    return e;
}
// one fibre per thread: 768 threads, or 384 with the register budget of two CTAs per SM
template <int MODE, bool FMA, int ROWS>
cudaError_t whole_launch_tg(int threads, bool timeline, const WholeArgs &a, const PackedTri<T> &M, int grid, size_t smem, bool set_attr, cudaStream_t s) {
#if LPFNT_MAXT == 32
    if constexpr (MODE == kLowerMatvec)
        if (timeline)
            return threads == 384 ? whole_launch<MODE, FMA, 384, ROWS, 1, false, 2, true>(a, M, grid, smem, set_attr, s) : whole_launch<MODE, FMA, 768, ROWS, 1, false, 1, true>(a, M, grid, smem, set_attr, s);
#endif
    return threads == 384 ? whole_launch<MODE, FMA, 384, ROWS, 1, false, 2>(a, M, grid, smem, set_attr, s) : whole_launch<MODE, FMA, 768, ROWS>(a, M, grid, smem, set_attr, s);
}
// two fibres per thread, 384 threads, one CTA per SM, side by side (lower mat-vec)
template <int MODE, bool FMA, bool SEQ>
cudaError_t whole_launch_two(bool timeline, const WholeArgs &a, const PackedTri<T> &M, int grid, size_t smem, bool set_attr, cudaStream_t s) {
#if LPFNT_MAXT == 32
    if constexpr (MODE == kLowerMatvec)
        if (timeline) return whole_launch<MODE, FMA, 384, 2, 2, SEQ, 1, true>(a, M, grid, smem, set_attr, s);
#endif
    return whole_launch<MODE, FMA, 384, 2, 2, SEQ, 1>(a, M, grid, smem, set_attr, s);
}
}  // namespace

// fibres: 1; 2 = two per thread side by side
template <>
cudaError_t launch_whole<T>(int mode, bool fma, int threads, int rows, int fibres, bool timeline, const WholeArgs &a, const double *tri, int grid, size_t smem, bool set_attr, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    if (threads != 384 && threads != 768) return cudaErrorInvalidValue;
    if (a.n_peers > 0) {
        // result pushed into every rank's symmetric buffer: the two default shapes of the lower mat-vec at the large capacities
#if LPFNT_MAXT >= 24
        if (mode == kLowerMatvec && fibres == 2 && threads == 384 && !timeline)
            return fma ? whole_launch<kLowerMatvec, true, 384, 2, 2, false, 1, false, true>(a, M, grid, smem, set_attr, s)
                       : whole_launch<kLowerMatvec, false, 384, 2, 2, false, 1, false, true>(a, M, grid, smem, set_attr, s);
        if (mode == kLowerMatvec && fibres == 1 && threads == 768 && !timeline)
            return fma ? whole_launch<kLowerMatvec, true, 768, 2, 1, false, 1, false, true>(a, M, grid, smem, set_attr, s)
                       : whole_launch<kLowerMatvec, false, 768, 2, 1, false, 1, false, true>(a, M, grid, smem, set_attr, s);
#if LPFNT_MAXT == 24
        if (mode == kLowerMatvec && fibres == 1 && threads == 384 && !timeline)
            return fma ? whole_launch<kLowerMatvec, true, 384, 2, 1, false, 2, false, true>(a, M, grid, smem, set_attr, s)
                       : whole_launch<kLowerMatvec, false, 384, 2, 1, false, 2, false, true>(a, M, grid, smem, set_attr, s);
#endif
#endif
        if (std::getenv("LPFNT_DEBUG_OCCUPANCY")) std::fprintf(stderr, "launch_whole<%d>: no push variant for mode %d fma %d threads %d fibres %d timeline %d\n", T, mode, int(fma), threads, fibres, int(timeline));
        return cudaErrorNotSupported;
    }
    if (fibres == 2) {
        if (mode != kLowerMatvec || threads != 384) return cudaErrorInvalidValue;
        return fma ? whole_launch_two<kLowerMatvec, true, false>(timeline, a, M, grid, smem, set_attr, s) : whole_launch_two<kLowerMatvec, false, false>(timeline, a, M, grid, smem, set_attr, s);
    }
    if (fibres != 1) return cudaErrorInvalidValue;  // (two fibres one after the other -- "fibres 3" -- lost to both other forms and was removed)
    (void)rows;  // every mode folds two rows side by side (four-row blocks lost to the two-fibre forms and were removed)
    switch (mode) {
        case kLowerMatvec:
            return fma ? whole_launch_tg<kLowerMatvec, true, 2>(threads, timeline, a, M, grid, smem, set_attr, s) : whole_launch_tg<kLowerMatvec, false, 2>(threads, timeline, a, M, grid, smem, set_attr, s);
        case kUpperMatvec: return fma ? whole_launch_tg<kUpperMatvec, true, 2>(threads, false, a, M, grid, smem, set_attr, s) : whole_launch_tg<kUpperMatvec, false, 2>(threads, false, a, M, grid, smem, set_attr, s);
        case kLowerSolve: return whole_launch_tg<kLowerSolve, false, 2>(threads, false, a, M, grid, smem, set_attr, s);
        default: return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t launch_whole_split<T>(bool fma, bool timeline, const WholeSplitArgs &a, const double *tri, int grid, size_t smem, bool set_attr, cudaStream_t s) {
#if LPFNT_MAXT == 32
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    auto go = [&](auto kern) -> cudaError_t {
        if (set_attr) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            // two CTAs per SM need the full shared-memory carve-out
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            if (std::getenv("LPFNT_DEBUG_OCCUPANCY")) {
                int nb = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kSplitThreads, smem);
                std::fprintf(stderr, "k_whole_split: %zu bytes of dynamic shared memory, %d CTA(s) per SM\n", smem, nb);
            }
        }
        kern<<<grid, kSplitThreads, smem, s>>>(a, M);
        return cudaGetLastError();
    };
#ifdef LPFNT_TIMELINE_SPLIT  // development build (LPFNT_TIMELINE_SPLIT=1 python -m lpfun_b200._build): phase stamps, tools/timeline.py
    if (timeline) return fma ? go(k_whole_split<T, true, true>) : go(k_whole_split<T, false, true>);
#endif
    (void)timeline;
    return fma ? go(k_whole_split<T, true>) : go(k_whole_split<T, false>);
#else
    (void)fma; (void)timeline; (void)a; (void)tri; (void)grid; (void)smem; (void)set_attr; (void)s;
    return cudaErrorInvalidValue;
#endif
}

template <>
cudaError_t launch_slab<T>(int mode, bool fma, const SlabArgs &a, const double *tri, int grid, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec:
            if (fma) k_slab<T, kLowerMatvec, true><<<grid, kSlabThreads, 0, s>>>(a, M);
            else k_slab<T, kLowerMatvec, false><<<grid, kSlabThreads, 0, s>>>(a, M);
            break;
        case kUpperMatvec:
            if (fma) k_slab<T, kUpperMatvec, true><<<grid, kSlabThreads, 0, s>>>(a, M);
            else k_slab<T, kUpperMatvec, false><<<grid, kSlabThreads, 0, s>>>(a, M);
            break;
        case kLowerSolve: k_slab<T, kLowerSolve, false><<<grid, kSlabThreads, 0, s>>>(a, M); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <>
cudaError_t launch_group_sweep<T>(int mode, bool fma, const GroupSweepArgs &a, const double *tri, int batch, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? group_launch<kLowerMatvec, true>(a, M, batch, s) : group_launch<kLowerMatvec, false>(a, M, batch, s);
        case kUpperMatvec: return fma ? group_launch<kUpperMatvec, true>(a, M, batch, s) : group_launch<kUpperMatvec, false>(a, M, batch, s);
        case kLowerSolve: return group_launch<kLowerSolve, false>(a, M, batch, s);
        case kUpperSolve: return group_launch<kUpperSolve, false>(a, M, batch, s);
        case kUpperSolveDesc: return group_launch<kUpperSolveDesc, false>(a, M, batch, s);
        case kUpperMatvecDesc: return group_launch<kUpperMatvecDesc, false>(a, M, batch, s);
    }
    return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_line_sweep<T>(int mode, bool fma, const LineSweepArgs &a, const double *tri, int batch, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? line_launch<kLowerMatvec, true>(a, M, batch, s) : line_launch<kLowerMatvec, false>(a, M, batch, s);
        case kUpperMatvec: return fma ? line_launch<kUpperMatvec, true>(a, M, batch, s) : line_launch<kUpperMatvec, false>(a, M, batch, s);
        case kLowerSolve: return line_launch<kLowerSolve, false>(a, M, batch, s);
        case kUpperSolve: return line_launch<kUpperSolve, false>(a, M, batch, s);
        case kUpperMatvecDesc: return line_launch<kUpperMatvecDesc, false>(a, M, batch, s);
    }
    return cudaErrorInvalidValue;
}

#if LPFNT_MAXT == 8
cudaError_t launch_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int batch, bool restrict_, cudaStream_t s) {
    dim3 grid(unsigned((n_small + 255) / 256), batch);
    k_embed<<<grid, 256, 0, s>>>(in, out, phi, n_small, n_big, restrict_ ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_push_rows(const PushArgs &a, int grid, cudaStream_t s) {
    static const int carve = [] {
        const char *v = std::getenv("LPFNT_DEBUG_PUSH_CARVEOUT");
        const int c = v ? atoi(v) : -1;
        if (c >= 0) cudaFuncSetAttribute(k_push_rows, cudaFuncAttributePreferredSharedMemoryCarveout, c);
        return c;
    }();
    (void)carve;
    k_push_rows<<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_permute(const double *x, double *y, const int32_t *perm, int64_t n, int64_t stride, int batch, bool scatter, cudaStream_t s) {
    dim3 grid(unsigned((n + 255) / 256), batch);
    k_permute<<<grid, 256, 0, s>>>(x, y, perm, n, stride, scatter ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t launch_generic_sweep(const GenericArgs &a, int batch, cudaStream_t s) {
    dim3 grid((a.num_threads + 127) / 128, batch);
    k_sweep_generic<<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}
#endif

}  // namespace lpfnt
