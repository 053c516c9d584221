#pragma once
#include <cuda_runtime.h>

#include "eval_kernels.cuh"

namespace lpfnt {
cudaError_t launch_eval_horner(const EvalArgs &a, const double *nodes, int n1, cudaStream_t s);
cudaError_t launch_eval_block(const EvalBlockArgs &a, const double *nodes, int n1, bool chebyshev, cudaStream_t s);
cudaError_t launch_eval_pad(const EvalPadArgs &a, cudaStream_t s);
// few points: subtree partial sums + basis-form combination (two launches)
cudaError_t launch_eval_split(const EvalSubArgs &a, const double *nodes, int n1, cudaStream_t s);
cudaError_t launch_eval_generic(const EvalGenericArgs &a, cudaStream_t s);
cudaError_t launch_fp64_probe(double *sink, int iters, int blocks, cudaStream_t s);
}  // namespace lpfnt
