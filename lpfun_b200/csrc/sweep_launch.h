// Launch entry points of the register-fibre sweep kernels; one translation unit per MAXT
// (sweep_inst.cu compiled with -DLPFNT_MAXT=8/16/24/32) so that nvcc compiles them in parallel.
#pragma once
#include <cuda_runtime.h>

#include "sweep_kernels.cuh"
#include "whole_kernel.cuh"
#include "whole_split_kernel.cuh"

namespace lpfnt {

constexpr int kSweepThreads = 128;

// tri: packed triangle in KERNEL layout (see PackedTri), MAXT*(MAXT+1)/2 doubles on the host.
template <int MAXT>
cudaError_t launch_line_sweep(int mode, bool fma, const LineSweepArgs &a, const double *tri, int batch, cudaStream_t s);
template <int MAXT>
cudaError_t launch_group_sweep(int mode, bool fma, const GroupSweepArgs &a, const double *tri, int batch, cudaStream_t s);
// slab kernel: coordinates 0 and 1 in one pass (any m >= 2)
template <int MAXT>
cudaError_t launch_slab(int mode, bool fma, const SlabArgs &a, const double *tri, int grid, cudaStream_t s);
// whole-function kernel (m = 2, 3): threads = 384 or 768; rows = 2 or 4 output rows folded side by side; fibres = 1 or 2
// items per thread (2: lower mat-vec, 384 threads); timeline: the instantiation with clock64 stamps (MAXT = 32, lower mat-vec only).
// set_attr: opt in to `smem` bytes of dynamic shared memory first (per device and kernel, tracked by the plan).
template <int MAXT>
cudaError_t launch_whole(int mode, bool fma, int threads, int rows, int fibres, bool timeline, const WholeArgs &a, const double *tri, int grid, size_t smem, bool set_attr, cudaStream_t s);

// split whole-function kernel (m = 3, lower mat-vec, capacities 24 and 32): two units per function, two CTAs per SM
template <int MAXT>
cudaError_t launch_whole_split(bool fma, bool timeline, const WholeSplitArgs &a, const double *tri, int grid, size_t smem, bool set_attr, cudaStream_t s);

cudaError_t launch_push_rows(const PushArgs &a, int grid, cudaStream_t s);
cudaError_t launch_permute(const double *x, double *y, const int32_t *perm, int64_t n, int64_t stride, int batch, bool scatter, cudaStream_t s);
cudaError_t launch_embed(const double *in, double *out, const int64_t *phi, int64_t n_small, int64_t n_big, int batch, bool restrict_, cudaStream_t s);
cudaError_t launch_generic_sweep(const GenericArgs &a, int batch, cudaStream_t s);

}  // namespace lpfnt
