// Launch entry points of the register-fibre sweep kernels; one translation unit per MAXT
// (sweep_inst.cu compiled with -DLPFNT_MAXT=8/16/24/32) so that nvcc compiles them in parallel.
#pragma once
#include <cuda_runtime.h>

#include "sweep_kernels.cuh"

namespace lpfnt {

constexpr int kSweepThreads = 128;

// tri: packed triangle in KERNEL layout (see PackedTri), MAXT*(MAXT+1)/2 doubles on the host.
template <int MAXT>
cudaError_t launch_line_sweep(int mode, bool fma, const LineSweepArgs &a, const double *tri, int batch, cudaStream_t s);
template <int MAXT>
cudaError_t launch_fibre_sweep(int mode, bool fma, const FibreSweepArgs &a, const double *tri, int batch, cudaStream_t s);

template <int MAXT>
cudaError_t launch_tile_sweep(int mode, bool fma, bool dim0, const TileSweepArgs &a, const double *tri, int max_line, int num_tiles, int batch, cudaStream_t s);

// dynamic shared memory of k_sweep_tiles: 2*threads rows of `pad` doubles
inline size_t tile_smem_bytes(int pad) { return size_t(2) * kSweepThreads * pad * sizeof(double); }

template <int MAXT>
cudaError_t launch_tile_pipe(int mode, bool fma, const PipeArgs &a, const double *tri, int grid, cudaStream_t s);

template <int MAXT>
cudaError_t launch_sorted_sweep(int mode, bool fma, bool dim0, const SortedSweepArgs &a, const double *tri, int threads, int batch, cudaStream_t s);

template <int MAXT>
cudaError_t launch_group_sweep(int mode, bool fma, const GroupSweepArgs &a, const double *tri, int batch, cudaStream_t s);

// whole-function kernel (m <= 3): dynamic shared memory = N doubles + per-coordinate item / offset tables
inline size_t whole_smem_bytes(int n_elem, int n_lines) {
    return size_t((n_elem + 1) & ~1) * sizeof(double) + size_t(3) * n_lines * sizeof(int2) + (size_t(32) + 2 * size_t(n_lines)) * sizeof(uint16_t) + 16;
}
template <int MAXT>
cudaError_t launch_whole(int mode, bool fma, const WholeArgs &a, const double *tri, int grid, cudaStream_t s);
// second generation (bulk-copy prefetch, one round per sweep, result stored from registers)
inline size_t whole2_smem_bytes(int n_elem, int n_lines, bool perm) {
    return size_t((n_elem + 3) & ~1) * sizeof(double) + 16 + size_t(3) * n_lines * sizeof(int2) + size_t((32 + 2 * n_lines + 7) & ~7) * sizeof(uint16_t) +
           (perm ? size_t((n_elem + 8) & ~7) * sizeof(uint16_t) : 0) + 16;
}
// slab kernel: coordinates 0 and 1 in one pass (any m >= 2)
template <int MAXT>
cudaError_t launch_slab(int mode, bool fma, const SlabArgs &a, const double *tri, int grid, cudaStream_t s);
// third generation (box layout, producer warps)
template <int MAXT>
cudaError_t launch_whole3(int mode, bool fma, const Whole3Args &a, const BoxConst &c, const double *tri, int grid, cudaStream_t s);
template <int MAXT>
cudaError_t launch_whole2(int mode, bool fma, const WholeArgs &a, const double *tri, int grid, cudaStream_t s);

template <int MAXT>
cudaError_t launch_plane_sweep(int mode, bool fma, const PlaneSweepArgs &a, const double *tri, int batch, cudaStream_t s);

cudaError_t launch_permute(const double *x, double *y, const int32_t *perm, int64_t n, int64_t stride, int batch, bool scatter, cudaStream_t s);
cudaError_t launch_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int batch, bool restrict_, cudaStream_t s);
cudaError_t launch_generic_sweep(const GenericArgs &a, int batch, cudaStream_t s);

}  // namespace lpfnt
