// Instantiations + launchers of the point-evaluation kernels and the FP64 probe.
#include <cstdlib>
#include <cstring>

#define LPFNT_EVAL_TU 1

#include "eval_launch.h"

namespace lpfnt {

namespace {
constexpr int kEvalThreads = 128;

template <int MAXT, int PPT>
cudaError_t eval_launch(const EvalArgs &a, const double *nodes, int n1, cudaStream_t s) {
    NodeTable<MAXT> nt;
    for (int k = 0; k < MAXT; ++k) nt.v[k] = (k < n1) ? nodes[k] : 0.0;
    const int64_t per_block = int64_t(kEvalThreads) * PPT;
    const int64_t blocks = (a.num_points + per_block - 1) / per_block;
    if (a.m <= 4) k_eval_horner<MAXT, PPT, 4, kEvalThreads><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    else if (a.m <= 8) k_eval_horner<MAXT, PPT, 8, kEvalThreads><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    else k_eval_horner<MAXT, PPT, kEvalMaxM, kEvalThreads><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_eval_horner(const EvalArgs &a, const double *nodes, int n1, cudaStream_t s) {
    if (n1 <= 8) return eval_launch<8, 2>(a, nodes, n1, s);
    if (n1 <= 16) return eval_launch<16, 2>(a, nodes, n1, s);
    if (n1 <= 24) return eval_launch<24, 2>(a, nodes, n1, s);
    if (n1 <= 32) return eval_launch<32, 1>(a, nodes, n1, s);
    return cudaErrorInvalidValue;
}

namespace {
template <int MAXT, int PPT, int MINB, bool CHEB = false>
cudaError_t eval_tree_launch(const EvalBlockArgs &a, const double *nodes, int n1, cudaStream_t s) {
    NodeTable<MAXT> nt;
    for (int k = 0; k < MAXT; ++k) nt.v[k] = (k < n1) ? nodes[k] : 0.0;
    const int64_t per_block = int64_t(kEvalThreads) * PPT;
    const int64_t blocks = (a.num_points + per_block - 1) / per_block;
    if (a.m == 1) k_eval_tree<MAXT, PPT, 1, kEvalThreads, MINB, CHEB><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    else if (a.m <= 4) k_eval_tree<MAXT, PPT, 4, kEvalThreads, MINB, CHEB><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    else if (a.m <= 8) k_eval_tree<MAXT, PPT, 8, kEvalThreads, MINB, CHEB><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    else k_eval_tree<MAXT, PPT, kEvalMaxM, kEvalThreads, MINB, CHEB><<<unsigned(blocks), kEvalThreads, 0, s>>>(a, nt);
    return cudaGetLastError();
}
}  // namespace

// many points, Newton or Chebyshev basis: k_eval_tree on the padded, block-staged coefficients
cudaError_t launch_eval_block(const EvalBlockArgs &a, const double *nodes, int n1, bool chebyshev, cudaStream_t s) {
    if (!a.tlines) return cudaErrorInvalidValue;
    if (chebyshev) {
        if (n1 <= 8) return eval_tree_launch<8, 2, 3, true>(a, nodes, n1, s);
        if (n1 <= 16) return eval_tree_launch<16, 2, 3, true>(a, nodes, n1, s);
        if (n1 <= 24) return eval_tree_launch<24, 2, 3, true>(a, nodes, n1, s);  // (204 registers / two CTAs per SM: 120 vs 113 ms)
        if (n1 <= 32) return eval_tree_launch<32, 1, 4, true>(a, nodes, n1, s);
        return cudaErrorInvalidValue;
    }
    if (n1 <= 8) return eval_tree_launch<8, 2, 3>(a, nodes, n1, s);
    if (n1 <= 16) return eval_tree_launch<16, 2, 3>(a, nodes, n1, s);
    if (n1 <= 24) return eval_tree_launch<24, 2, 3>(a, nodes, n1, s);
    if (n1 <= 32) return eval_tree_launch<32, 1, 4>(a, nodes, n1, s);  // (two points per thread at 198 registers: no faster)
    return cudaErrorInvalidValue;
}

namespace {
template <int MAXT>
cudaError_t eval_split_launch(const EvalSubArgs &a, const double *nodes, int n1, cudaStream_t s) {
    NodeTable<MAXT> nt;
    for (int k = 0; k < MAXT; ++k) nt.v[k] = (k < n1) ? nodes[k] : 0.0;
    const int64_t pairs = int64_t(a.num_sub) * a.num_points;
    k_eval_sub<MAXT><<<unsigned((pairs + 127) / 128), 128, 0, s>>>(a, nt);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_eval_combine<MAXT><<<unsigned(a.num_points), 256, 0, s>>>(a, nt);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_eval_split(const EvalSubArgs &a, const double *nodes, int n1, cudaStream_t s) {
    if (n1 <= 8) return eval_split_launch<8>(a, nodes, n1, s);
    if (n1 <= 16) return eval_split_launch<16>(a, nodes, n1, s);
    if (n1 <= 24) return eval_split_launch<24>(a, nodes, n1, s);
    if (n1 <= 32) return eval_split_launch<32>(a, nodes, n1, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_eval_pad(const EvalPadArgs &a, cudaStream_t s) {
    k_eval_pad<<<(a.num_lines + 3) / 4, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_eval_generic(const EvalGenericArgs &a, cudaStream_t s) {
    const int64_t blocks = (a.num_points + 127) / 128;
    k_eval_generic<<<unsigned(blocks), 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fp64_probe(double *sink, int iters, int blocks, cudaStream_t s) {
    k_fp64_probe<<<blocks, 256, 0, s>>>(sink, iters, 0.5);
    return cudaGetLastError();
}

}  // namespace lpfnt
