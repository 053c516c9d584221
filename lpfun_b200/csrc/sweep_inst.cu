// Instantiates the register-fibre sweep kernels for one fibre capacity LPFNT_MAXT.
#include <cstring>

#include "sweep_launch.h"

#ifndef LPFNT_MAXT
#error "compile with -DLPFNT_MAXT=<8|16|24|32>"
#endif

namespace lpfnt {

namespace {
constexpr int T = LPFNT_MAXT;
// shared-memory row pitches instantiated for this capacity: the generic odd pitch T+1 and a
// tight one for the common degrees (n = 30 -> 31, n = 20 -> 21, n = 10 -> 11)
constexpr int kPadWide = T + 1;
constexpr int kPadTight = (T == 32) ? 31 : (T == 24) ? 21 : (T == 16) ? 11 : 7;

template <int MODE, bool FMA>
cudaError_t line_launch(const LineSweepArgs &a, const PackedTri<T> &M, int batch, cudaStream_t s) {
    dim3 grid((a.num_lines + kSweepThreads - 1) / kSweepThreads, batch);
    k_sweep_lines<T, MODE, FMA, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    return cudaGetLastError();
}
template <int MODE, bool FMA>
cudaError_t fibre_launch(const FibreSweepArgs &a, const PackedTri<T> &M, int batch, cudaStream_t s) {
    dim3 grid((a.num_threads + kSweepThreads - 1) / kSweepThreads, batch);
    k_sweep_fibres<T, MODE, FMA, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    return cudaGetLastError();
}
template <int MODE, bool FMA>
cudaError_t pipe_launch(const PipeArgs &a, const PackedTri<T> &M, int grid, cudaStream_t s) {
    auto kern = k_tiles_pipe<T, MODE, FMA, kSweepThreads>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPipeSmemBytes));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, kSweepThreads, kPipeSmemBytes, s>>>(a, M);
    return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t launch_tile_pipe<T>(int mode, bool fma, const PipeArgs &a, const double *tri, int grid, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? pipe_launch<kLowerMatvec, true>(a, M, grid, s) : pipe_launch<kLowerMatvec, false>(a, M, grid, s);
        case kUpperMatvec: return fma ? pipe_launch<kUpperMatvec, true>(a, M, grid, s) : pipe_launch<kUpperMatvec, false>(a, M, grid, s);
        case kLowerSolve: return pipe_launch<kLowerSolve, false>(a, M, grid, s);
        case kUpperSolve: return pipe_launch<kUpperSolve, false>(a, M, grid, s);
    }
    return cudaErrorInvalidValue;
}

template <int MODE, bool FMA>
cudaError_t tile_launch(bool dim0, const TileSweepArgs &a, const PackedTri<T> &M, int num_tiles, int batch, cudaStream_t s) {
    dim3 grid(num_tiles, batch);
    if (dim0) k_sweep_tiles<T, MODE, FMA, true, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    else k_sweep_tiles<T, MODE, FMA, false, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    return cudaGetLastError();
}

namespace {
template <int MODE, bool FMA, bool DIM0>
cudaError_t sorted_launch(const SortedSweepArgs &a, const PackedTri<T> &M, int threads, int batch, cudaStream_t s) {
    if (threads > 256) {
        dim3 grid((a.num_items + 767) / 768, batch);
        k_sweep_sorted<T, MODE, FMA, DIM0, 768><<<grid, 768, 0, s>>>(a, M);
    } else {
        dim3 grid((a.num_items + 127) / 128, batch);
        k_sweep_sorted<T, MODE, FMA, DIM0, 128><<<grid, 128, 0, s>>>(a, M);
    }
    return cudaGetLastError();
}
template <int MODE, bool FMA>
cudaError_t sorted_launch2(bool dim0, const SortedSweepArgs &a, const PackedTri<T> &M, int threads, int batch, cudaStream_t s) {
    return dim0 ? sorted_launch<MODE, FMA, true>(a, M, threads, batch, s) : sorted_launch<MODE, FMA, false>(a, M, threads, batch, s);
}
}  // namespace

template <>
cudaError_t launch_sorted_sweep<T>(int mode, bool fma, bool dim0, const SortedSweepArgs &a, const double *tri, int threads, int batch, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? sorted_launch2<kLowerMatvec, true>(dim0, a, M, threads, batch, s) : sorted_launch2<kLowerMatvec, false>(dim0, a, M, threads, batch, s);
        case kUpperMatvec: return fma ? sorted_launch2<kUpperMatvec, true>(dim0, a, M, threads, batch, s) : sorted_launch2<kUpperMatvec, false>(dim0, a, M, threads, batch, s);
        case kLowerSolve: return sorted_launch2<kLowerSolve, false>(dim0, a, M, threads, batch, s);
        case kUpperSolve: return sorted_launch2<kUpperSolve, false>(dim0, a, M, threads, batch, s);
    }
    return cudaErrorInvalidValue;
}

namespace {
template <int MODE, bool FMA>
cudaError_t group_launch(const GroupSweepArgs &a, const PackedTri<T> &M, int batch, cudaStream_t s) {
    if (a.two_lanes && (MODE == kLowerMatvec || MODE == kUpperMatvec)) {
        dim3 grid((4 * int64_t(a.num_groups) + kSweepThreads - 1) / kSweepThreads, batch);
        k_sweep_groups2<T, MODE, FMA, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
        return cudaGetLastError();
    }
    dim3 grid((8 * int64_t(a.num_groups) + kSweepThreads - 1) / kSweepThreads, batch);
    k_sweep_groups<T, MODE, FMA, kSweepThreads><<<grid, kSweepThreads, 0, s>>>(a, M);
    return cudaGetLastError();
}
}  // namespace

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
    }
    return cudaErrorInvalidValue;
}

namespace {
template <int MODE, bool FMA, int ITEMS, int THREADS>
cudaError_t whole_launch(const WholeArgs &a, const PackedTri<T> &M, int grid, cudaStream_t s) {
    auto kern = k_whole<T, MODE, FMA, ITEMS, THREADS>;
    const size_t smem = whole_smem_bytes(a.n_elem, a.n_lines);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, THREADS, smem, s>>>(a, M);
    return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t launch_whole<T>(int mode, bool fma, const WholeArgs &a, const double *tri, int grid, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    // exact arithmetic: one fibre per thread and many warps; FMA: two fibres per thread share the constants
    switch (mode) {
        case kLowerMatvec: return fma ? whole_launch<kLowerMatvec, true, 2, 256>(a, M, grid, s) : whole_launch<kLowerMatvec, false, 1, 384>(a, M, grid, s);
        case kLowerSolve: return whole_launch<kLowerSolve, false, 1, 384>(a, M, grid, s);
        default: return cudaErrorInvalidValue;
    }
}

namespace {
template <int MODE, bool FMA>
cudaError_t whole2_launch(const WholeArgs &a, const PackedTri<T> &M, int grid, cudaStream_t s) {
    auto kern = k_whole2<T, MODE, FMA, kWhole2Threads>;
    const size_t smem = whole2_smem_bytes(a.n_elem, a.n_lines, a.perm_in || a.perm_out);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, kWhole2Threads, smem, s>>>(a, M);
    return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t launch_whole2<T>(int mode, bool fma, const WholeArgs &a, const double *tri, int grid, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? whole2_launch<kLowerMatvec, true>(a, M, grid, s) : whole2_launch<kLowerMatvec, false>(a, M, grid, s);
        case kUpperMatvec: return fma ? whole2_launch<kUpperMatvec, true>(a, M, grid, s) : whole2_launch<kUpperMatvec, false>(a, M, grid, s);
        case kLowerSolve: return whole2_launch<kLowerSolve, false>(a, M, grid, s);
        default: return cudaErrorInvalidValue;
    }
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

namespace {
template <int MODE, bool FMA>
cudaError_t whole3_launch(const Whole3Args &a, const BoxConst &c, const PackedTri<T> &M, int grid, cudaStream_t s) {
    auto kern = k_whole3<T, MODE, FMA>;
    const size_t smem = whole3_smem_bytes(a.n_box, a.n_lines, a.n_elem, a.perm16 != nullptr);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    kern<<<grid, kW3Threads, smem, s>>>(a, M, c);
    return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t launch_whole3<T>(int mode, bool fma, const Whole3Args &a, const BoxConst &c, const double *tri, int grid, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? whole3_launch<kLowerMatvec, true>(a, c, M, grid, s) : whole3_launch<kLowerMatvec, false>(a, c, M, grid, s);
        case kUpperMatvec: return fma ? whole3_launch<kUpperMatvec, true>(a, c, M, grid, s) : whole3_launch<kUpperMatvec, false>(a, c, M, grid, s);
        case kLowerSolve: return whole3_launch<kLowerSolve, false>(a, c, M, grid, s);
        default: return cudaErrorInvalidValue;
    }
}

namespace {
template <int MODE, bool FMA>
cudaError_t plane_launch(const PlaneSweepArgs &a, const PackedTri<T> &M, int batch, cudaStream_t s) {
    dim3 grid((a.num_planes + 3) / 4, batch);
    k_sweep_planes<T, MODE, FMA><<<grid, 128, 0, s>>>(a, M);
    return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t launch_plane_sweep<T>(int mode, bool fma, const PlaneSweepArgs &a, const double *tri, int batch, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? plane_launch<kLowerMatvec, true>(a, M, batch, s) : plane_launch<kLowerMatvec, false>(a, M, batch, s);
        case kUpperMatvec: return fma ? plane_launch<kUpperMatvec, true>(a, M, batch, s) : plane_launch<kUpperMatvec, false>(a, M, batch, s);
        case kLowerSolve: return plane_launch<kLowerSolve, false>(a, M, batch, s);
        case kUpperSolve: return plane_launch<kUpperSolve, false>(a, M, batch, s);
    }
    return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_tile_sweep<T>(int mode, bool fma, bool dim0, const TileSweepArgs &a, const double *tri, int max_line, int num_tiles, int batch, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? tile_launch<kLowerMatvec, true>(dim0, a, M, num_tiles, batch, s) : tile_launch<kLowerMatvec, false>(dim0, a, M, num_tiles, batch, s);
        case kUpperMatvec: return fma ? tile_launch<kUpperMatvec, true>(dim0, a, M, num_tiles, batch, s) : tile_launch<kUpperMatvec, false>(dim0, a, M, num_tiles, batch, s);
        case kLowerSolve: return tile_launch<kLowerSolve, false>(dim0, a, M, num_tiles, batch, s);
        case kUpperSolve: return tile_launch<kUpperSolve, false>(dim0, a, M, num_tiles, batch, s);
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
    }
    return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_fibre_sweep<T>(int mode, bool fma, const FibreSweepArgs &a, const double *tri, int batch, cudaStream_t s) {
    PackedTri<T> M;
    std::memcpy(M.v, tri, sizeof(M.v));
    switch (mode) {
        case kLowerMatvec: return fma ? fibre_launch<kLowerMatvec, true>(a, M, batch, s) : fibre_launch<kLowerMatvec, false>(a, M, batch, s);
        case kUpperMatvec: return fma ? fibre_launch<kUpperMatvec, true>(a, M, batch, s) : fibre_launch<kUpperMatvec, false>(a, M, batch, s);
        case kLowerSolve: return fibre_launch<kLowerSolve, false>(a, M, batch, s);
        case kUpperSolve: return fibre_launch<kUpperSolve, false>(a, M, batch, s);
        case kUpperSolveDesc: return fibre_launch<kUpperSolveDesc, false>(a, M, batch, s);
    }
    return cudaErrorInvalidValue;
}

#if LPFNT_MAXT == 8
cudaError_t launch_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int batch, bool restrict_, cudaStream_t s) {
    dim3 grid(unsigned((n_small + 255) / 256), batch);
    k_embed<<<grid, 256, 0, s>>>(in, out, phi, n_small, n_big, restrict_ ? 1 : 0);
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
