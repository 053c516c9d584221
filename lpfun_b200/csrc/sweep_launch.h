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
cudaError_t launch_tile_sweep(int mode, bool fma, bool dim0, const TileSweepArgs &a, const double *tri, int num_tiles, int batch, cudaStream_t s);

cudaError_t launch_generic_sweep(const GenericArgs &a, int batch, cudaStream_t s);

}  // namespace lpfnt
