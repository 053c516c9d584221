// C ABI of liblpfnt.so: plan management, operator dispatch, host<->device staging.
// See include/lpfnt.h for the contract of every entry point and the reference call sites
// (src/lpfun/transform.py, src/lpfun/core/molecules.py) each one replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "eval_launch.h"
#include "lpfnt_internal.h"
#include "sweep_launch.h"

using namespace lpfnt;

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                    \
            return LPFNT_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

namespace {

struct DeviceDim {
    int32_t *thr_fib = nullptr, *thr_start = nullptr, *fib_ptr = nullptr;
    int2 *ent = nullptr;
    int32_t *tile_fib = nullptr;
    uint32_t *items = nullptr, *rows0 = nullptr;
    int4 *tile_rec = nullptr;
    int2 *sorted = nullptr, *groups = nullptr;
    uint8_t *group_t = nullptr, *group_cta_cap = nullptr;
    uint16_t *ent_off16 = nullptr;           // line offsets in fibre order as uint16 (whole-function kernel, N < 65536)
    int32_t num_sorted = 0, num_groups = 0;
    int2 *tile_elem = nullptr;
    uint16_t *row_pos = nullptr;
    int2 *row_desc = nullptr;
    uint8_t *elem_row = nullptr;
    int32_t num_threads = 0, max_lines = 0, num_tiles = 0;
};

int tri_size(int n1) { return n1 * (n1 + 1) / 2; }

}  // namespace

struct lpfnt_plan {
    int device = 0;
    int m = 0, n = 0;
    int64_t N = 0, N1 = 0;
    int max_line = 0;
    bool has_perm = false;
    // device tables
    int32_t *d_line_off = nullptr, *d_perm = nullptr;
    int2 *d_sorted0 = nullptr;
    int32_t *d_line_order = nullptr;
    int32_t *d_plane_line = nullptr;         // planes + 1: first line of every plane (level-2 node)
    int32_t num_planes = 0;
    int use_planes = 0;                      // coordinates 0 and 1 fused, one warp per plane (selectable; measured slower
                                             // than the two separate passes: 0.61-0.66 ms vs 0.52 ms at d=3 n=30 B=4096)
    uint8_t *d_line_alpha = nullptr;
    int2 *d_eval_line = nullptr;
    int4 *d_eval_block = nullptr;
    double *d_cpad = nullptr;       // padded coefficient layout of the eval kernel
    int32_t num_eval_blocks = 0;
    int use_eval_block = 1;
    int64_t *d_A = nullptr;       // only for the generic eval fallback (uploaded lazily)
    double *d_nodes = nullptr, *d_mat = nullptr;  // generic-kernel matrix scratch
    DeviceDim dd[kMaxDim + 1];
    std::vector<int64_t> hA;      // kept only when the generic eval may need it
    // host copies of the 1-D operators (reference packing)
    std::vector<double> nodes, Vx_lt, inv_Vx_lt, Dx_ut[3], DxT_lt[3];
    // LU-factorised (non-triangular) V, e.g. the Chebyshev basis: upper factors; chebyshev_eval selects the
    // T_k basis in eval (transform.py:292-302, utils.py:165-195)
    std::vector<double> Vx_ut, inv_Vx_ut;
    int chebyshev_eval = 0;
    // workspaces (grown on demand): two lanes x {staged input, staged output, permutation scratch}
    double *ws[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t ws_bytes[6] = {0, 0, 0, 0, 0, 0};
    cudaStream_t stream = nullptr, stream2 = nullptr;  // used by the host-pointer paths (two lanes)
    int64_t launches = 0;
    int64_t table_bytes = 0;
    int force_generic = 0;
    int use_tiles = 0;                       // row-tile kernels (sorted work items, fused coordinate 0); measured slower than the
                                             // per-coordinate kernels on B200 so far, kept selectable
    int debug_skip = 0;
    int use_whole = -1;                      // whole function in shared memory, ONE pass over HBM per transform (m = 2, 3):
                                             // -1 auto (k_whole2 when its 768 threads are at least 2/3 busy, N_1 >= 512, and batch >= 8),
                                             // 0 off, 1 k_whole, 2 k_whole2, 3 k_whole3 (box layout, producer warp)
    // slab kernel (coordinates 0 and 1 in one pass, any m >= 2; see k_slab)
    struct Slabs {
        bool ok = false;
        int32_t n_slabs = 0, max_t = 0;
        SlabDesc *d_desc = nullptr;
        int2 *d_item0 = nullptr, *d_item1 = nullptr;
    } slabs;
    int use_slab = 1;
    int use_cta_cap = 1;                     // k_sweep_groups: per-CTA row capacity (8 / 16 / MAXT)
    // k_whole2 item orders per coordinate: [h][v], v = 0 sorted by length, 1 / 2 = length classes of 8 / 16 (natural order
    // inside a class), 3 = natural; w2_order_* = variant per coordinate (2 bits each) for the exact / FMA forms.
    // Strict length order keeps 97 % of the lanes busy but scatters a warp's lanes over many lines (shared-memory bank
    // conflicts, 2.5-3.6x the store sectors); natural order does the opposite.  Measured at d=3, n=30, B=4096 (tools/
    // order_sweep.py, all 64 combinations): classes of 16 for coordinates 1 and 2 are best for both forms (exact 0.619 vs
    // 0.663 ms in strict order, FMA with its scattered stores 0.463 vs 0.586 ms; classes of 8 are within 0.5 %);
    // coordinate 0 is best in strict order.
    int2 *w2_items[3][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    int w2_order_exact = 0 | (2 << 2) | (2 << 4), w2_order_fma = 0 | (2 << 2) | (2 << 4);
    int64_t host_chunk_bytes = 0;            // host-pointer calls: bytes per pipelined chunk (H2D | sweeps | D2H on two lanes); 0 = automatic
    long long *d_timeline = nullptr;         // development: clock64 stamps of k_whole2 (option "timeline")
    // box layout of the third-generation whole-function kernel (m = 2, 3; see k_whole3)
    struct Box {
        bool ok = false;
        BoxConst c{};
        int32_t n_box = 0, n_item[3] = {0, 0, 0}, max_t = 0;
        uint32_t *d_item[3] = {nullptr, nullptr, nullptr};
        int2 *d_pline = nullptr;
        uint16_t *d_perm16 = nullptr, *d_line_off16 = nullptr;
    } box;
    int use_two_lanes = 0;                   // two adjacent lanes per thread in the lane-group kernel
    int use_groups = 1;                      // height-ordered lane groups in the per-coordinate fibre kernel
    int use_sorted = 0;                      // globally length-sorted per-coordinate kernels (no shared memory)
    int use_pipe = 0;                        // persistent cp.async-pipelined variant of the row-tile kernels
    int num_sms = 148;
    int64_t l2_chunk_bytes = 0;              // > 0: batched multi-pass transforms run in chunks of this many bytes (L2 reuse)
    // optional per-launch tracing (CUDA events around every kernel launch)
    int trace = 0;
    struct TraceRec { int label; cudaEvent_t a, b; };
    std::vector<TraceRec> trace_recs;
};

namespace {

template <typename T>
int upload(T **dst, const T *src, size_t count, lpfnt_plan *p) {
    *dst = nullptr;
    if (count == 0) return LPFNT_OK;
    CK(cudaMalloc(reinterpret_cast<void **>(dst), count * sizeof(T)));
    CK(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    p->table_bytes += int64_t(count * sizeof(T));
    return LPFNT_OK;
}

int ensure_ws(lpfnt_plan *p, int slot, size_t bytes) {
    if (p->ws_bytes[slot] >= bytes) return LPFNT_OK;
    if (p->ws[slot]) CK(cudaFree(p->ws[slot]));
    p->ws[slot] = nullptr;
    p->ws_bytes[slot] = 0;
    CK(cudaMalloc(reinterpret_cast<void **>(&p->ws[slot]), bytes));
    p->ws_bytes[slot] = bytes;
    return LPFNT_OK;
}


// Tracing: label = 100 * kernel kind (0 line sweep, 1 fibre sweep, 2 generic sweep, 3 eval, 4 fused tile) + coordinate.
void trace_begin(lpfnt_plan *p, int label, cudaStream_t s) {
    if (!p->trace) return;
    lpfnt_plan::TraceRec r{label, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    p->trace_recs.push_back(r);
}
void trace_end(lpfnt_plan *p, cudaStream_t s) {
    if (!p->trace || p->trace_recs.empty()) return;
    cudaEventRecord(p->trace_recs.back().b, s);
}

// Re-pack a reference-layout triangle of size n1 into the kernels' compile-time layout of
// capacity cap (see PackedTri): lower stays row-major, upper becomes its transposed lower form.
void repack(const double *src, int n1, int kind, int cap, std::vector<double> &dst) {
    dst.assign(tri_size(cap), 0.0);
    // a sub-set of the index set (a slab of planes, sharding.py) may only have lines shorter than n1: its kernels use
    // the leading cap x cap block (atoms.py:1241-1243)
    const int nn = std::min(n1, cap);
    if (kind == LPFNT_LOWER) {
        std::copy(src, src + tri_size(nn), dst.begin());
    } else {
        for (int k = 0; k < nn; ++k) {
            const double *row = src + int64_t(k) * n1 - int64_t(k) * (k - 1) / 2;
            for (int j = k; j < nn; ++j) dst[j * (j + 1) / 2 + k] = row[j - k];
        }
    }
    // unit diagonal beyond n1 keeps the (never executed) padded solve rows finite
    for (int k = n1; k < cap; ++k) dst[k * (k + 1) / 2 + k] = 1.0;
}

int pick_cap(int len) {
    if (len <= 8) return 8;
    if (len <= 16) return 16;
    if (len <= 24) return 24;
    if (len <= 32) return 32;
    return 0;
}

struct SweepSpec {
    const double *packed;  // host, reference packing
    int kind;              // LPFNT_LOWER / LPFNT_UPPER
    bool solve;
    bool fma;
};

int sweep_mode(const SweepSpec &s) {
    if (s.kind == LPFNT_LOWER) return s.solve ? kLowerSolve : kLowerMatvec;
    return s.solve ? kUpperSolve : kUpperMatvec;
}

int launch_slab_pass(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s);

// One sweep along coordinate h: src -> dst (device pointers, may alias when no perm is used).
int launch_dim(lpfnt_plan *p, const SweepSpec &spec, int h, bool fuse0, const double *src, double *dst, int64_t batch,
               bool gather, bool scatter, cudaStream_t s) {
    const int n1 = p->n + 1;
    int mode = sweep_mode(spec);
    if (mode == kUpperSolve && h >= 1) mode = kUpperSolveDesc;  // the reference's fold order there (atoms.py:391-402)
    const int len = (h == 0) ? p->max_line : std::max(p->dd[h].max_lines, fuse0 ? p->max_line : 0);
    const bool slab_ok = p->use_slab && p->slabs.ok && !p->force_generic && !p->use_tiles && !p->use_planes && !p->use_sorted &&
                         pick_cap(std::max(p->slabs.max_t, 1)) > 0 && !(spec.solve && spec.kind == LPFNT_UPPER);
    if (fuse0 && h == 1 && slab_ok) return launch_slab_pass(p, spec, src, dst, batch, gather, scatter, s);
    const bool planes = fuse0 && h == 1 && p->use_planes && p->num_planes > 0;
    const bool tiles = !planes && h >= 1 && p->use_tiles && !p->use_sorted && p->dd[h].num_tiles > 0;
    if (fuse0 && !tiles && !planes) { set_error("internal: fused coordinate-0 sweep needs the tile kernel"); return LPFNT_EINVAL; }
    const int cap = p->force_generic ? 0 : pick_cap(std::max(len, 1));
    std::vector<double> tri;
    if (cap) repack(spec.packed, n1, spec.kind, cap, tri);
    else {
        // generic kernel reads the matrix from device memory (stream ordered copy of a small table)
        CK(cudaMemcpyAsync(p->d_mat, spec.packed, sizeof(double) * tri_size(n1), cudaMemcpyHostToDevice, s));
    }
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = int(std::min<int64_t>(65535, batch - b0));
        const double *in = src + b0 * p->N;
        double *out = dst + b0 * p->N;
        cudaError_t e = cudaSuccess;
        trace_begin(p, (cap == 0 ? 200 : (planes ? 1300 : (p->use_sorted ? 900 : (h == 0 ? 0 : (tiles ? (fuse0 ? 500 : 400) : 100))))) + h, s);
        if (planes && cap != 0) {
            PlaneSweepArgs pa{in, out, p->d_line_off, p->d_plane_line, gather ? p->d_perm : nullptr, scatter ? p->d_perm : nullptr, p->N, p->num_planes};
            switch (cap) {
                case 8: e = launch_plane_sweep<8>(mode, spec.fma, pa, tri.data(), nb, s); break;
                case 16: e = launch_plane_sweep<16>(mode, spec.fma, pa, tri.data(), nb, s); break;
                case 24: e = launch_plane_sweep<24>(mode, spec.fma, pa, tri.data(), nb, s); break;
                default: e = launch_plane_sweep<32>(mode, spec.fma, pa, tri.data(), nb, s); break;
            }
            trace_end(p, s);
            p->launches += 1;
            if (e != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
            continue;
        }
        const bool sorted_ok = cap != 0 && p->use_sorted && !fuse0 && (h == 0 ? p->d_sorted0 != nullptr : p->dd[h].sorted != nullptr);
        if (sorted_ok) {
            const int nit = (h == 0) ? int(p->N1) : p->dd[h].num_sorted;
            SortedSweepArgs a{in, out, h == 0 ? p->d_sorted0 : p->dd[h].sorted, h == 0 ? nullptr : p->dd[h].ent,
                              gather ? p->d_perm : nullptr, scatter ? p->d_perm : nullptr, p->N, nit};
            // batched: one CTA per function when it fits, so that the L1 keeps the function resident
            const int threads = (batch > 1 && nit <= 768) ? 768 : 128;
            switch (cap) {
                case 8: e = launch_sorted_sweep<8>(mode, spec.fma, h == 0, a, tri.data(), threads, nb, s); break;
                case 16: e = launch_sorted_sweep<16>(mode, spec.fma, h == 0, a, tri.data(), threads, nb, s); break;
                case 24: e = launch_sorted_sweep<24>(mode, spec.fma, h == 0, a, tri.data(), threads, nb, s); break;
                default: e = launch_sorted_sweep<32>(mode, spec.fma, h == 0, a, tri.data(), threads, nb, s); break;
            }
        } else if (cap == 0) {
            GenericArgs a{};
            a.in = in; a.out = out; a.mat = p->d_mat; a.line_off = p->d_line_off;
            a.perm_in = gather ? p->d_perm : nullptr;
            a.perm_out = scatter ? p->d_perm : nullptr;
            a.stride = p->N; a.n1 = n1; a.mode = mode; a.fma = spec.fma ? 1 : 0;
            if (h == 0) { a.ent = nullptr; a.num_threads = int32_t(p->N1); }
            else {
                const DeviceDim &d = p->dd[h];
                a.thr_fib = d.thr_fib; a.thr_start = d.thr_start; a.fib_ptr = d.fib_ptr; a.ent = d.ent;
                a.num_threads = d.num_threads;
            }
            e = launch_generic_sweep(a, nb, s);
        } else if (h == 0) {
            LineSweepArgs a{in, out, p->d_line_off, gather ? p->d_perm : nullptr, scatter ? p->d_perm : nullptr, p->d_line_order, p->N, int32_t(p->N1)};
            switch (cap) {
                case 8: e = launch_line_sweep<8>(mode, spec.fma, a, tri.data(), nb, s); break;
                case 16: e = launch_line_sweep<16>(mode, spec.fma, a, tri.data(), nb, s); break;
                case 24: e = launch_line_sweep<24>(mode, spec.fma, a, tri.data(), nb, s); break;
                default: e = launch_line_sweep<32>(mode, spec.fma, a, tri.data(), nb, s); break;
            }
        } else if (tiles) {
            const DeviceDim &d = p->dd[h];
            TileSweepArgs a{in, out, d.tile_rec, d.tile_elem, d.ent, d.row_pos, d.elem_row, d.items, d.rows0,
                            gather ? p->d_perm : nullptr, scatter ? p->d_perm : nullptr, p->N, p->debug_skip};
            switch (cap) {
                case 8: e = launch_tile_sweep<8>(mode, spec.fma, fuse0, a, tri.data(), p->max_line, d.num_tiles, nb, s); break;
                case 16: e = launch_tile_sweep<16>(mode, spec.fma, fuse0, a, tri.data(), p->max_line, d.num_tiles, nb, s); break;
                case 24: e = launch_tile_sweep<24>(mode, spec.fma, fuse0, a, tri.data(), p->max_line, d.num_tiles, nb, s); break;
                default: e = launch_tile_sweep<32>(mode, spec.fma, fuse0, a, tri.data(), p->max_line, d.num_tiles, nb, s); break;
            }
        } else {
            if (gather) { set_error("internal: gather is only fused into the coordinate-0 sweep"); return LPFNT_EINVAL; }
            const DeviceDim &d = p->dd[h];
            if (p->use_groups && d.groups) {
                GroupSweepArgs ga{in, out, d.groups, d.group_t, d.ent, scatter ? p->d_perm : nullptr, p->N, d.num_groups, p->use_two_lanes, p->use_cta_cap ? d.group_cta_cap : nullptr};
                switch (cap) {
                    case 8: e = launch_group_sweep<8>(mode, spec.fma, ga, tri.data(), nb, s); break;
                    case 16: e = launch_group_sweep<16>(mode, spec.fma, ga, tri.data(), nb, s); break;
                    case 24: e = launch_group_sweep<24>(mode, spec.fma, ga, tri.data(), nb, s); break;
                    default: e = launch_group_sweep<32>(mode, spec.fma, ga, tri.data(), nb, s); break;
                }
                trace_end(p, s);
                p->launches += 1;
                if (e != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
                continue;
            }
            FibreSweepArgs a{in, out, d.thr_fib, d.thr_start, d.fib_ptr, d.ent, scatter ? p->d_perm : nullptr, p->N, d.num_threads};
            switch (cap) {
                case 8: e = launch_fibre_sweep<8>(mode, spec.fma, a, tri.data(), nb, s); break;
                case 16: e = launch_fibre_sweep<16>(mode, spec.fma, a, tri.data(), nb, s); break;
                case 24: e = launch_fibre_sweep<24>(mode, spec.fma, a, tri.data(), nb, s); break;
                default: e = launch_fibre_sweep<32>(mode, spec.fma, a, tri.data(), nb, s); break;
            }
        }
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

// One pipelined row-tile pass along coordinate h (optionally fused with coordinate 0): src -> dst.
int launch_pipe(lpfnt_plan *p, const SweepSpec &spec, int h, bool fuse0, const double *src, double *dst, int64_t batch, cudaStream_t s) {
    const int n1 = p->n + 1;
    const int mode = sweep_mode(spec);
    const DeviceDim &d = p->dd[h];
    const int len = std::max(d.max_lines, fuse0 ? p->max_line : 0);
    const int cap = pick_cap(std::max(len, 1));
    std::vector<double> tri;
    repack(spec.packed, n1, spec.kind, cap, tri);
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 20)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 20, batch - b0));
        const int64_t units = int64_t(d.num_tiles) * nb;
        int upc = int((units + 2 * p->num_sms - 1) / (2 * p->num_sms));
        upc = std::max(1, std::min(upc, kPipeMaxUnits));
        const int grid = int((units + upc - 1) / upc);
        PipeArgs a{src + b0 * p->N, dst + b0 * p->N, d.tile_rec, d.tile_elem, d.row_desc, d.elem_row, d.items, d.rows0,
                   p->N, d.num_tiles, nb, fuse0 ? 1 : 0, upc, p->debug_skip};
        cudaError_t e = cudaSuccess;
        trace_begin(p, (fuse0 ? 700 : 600) + h, s);
        switch (cap) {
            case 8: e = launch_tile_pipe<8>(mode, spec.fma, a, tri.data(), grid, s); break;
            case 16: e = launch_tile_pipe<16>(mode, spec.fma, a, tri.data(), grid, s); break;
            case 24: e = launch_tile_pipe<24>(mode, spec.fma, a, tri.data(), grid, s); break;
            default: e = launch_tile_pipe<32>(mode, spec.fma, a, tri.data(), grid, s); break;
        }
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("pipe kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

int launch_perm(lpfnt_plan *p, const double *src, double *dst, int64_t batch, bool scatter, cudaStream_t s) {
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = int(std::min<int64_t>(65535, batch - b0));
        trace_begin(p, scatter ? 801 : 800, s);
        cudaError_t e = launch_permute(src + b0 * p->N, dst + b0 * p->N, p->d_perm, p->N, p->N, nb, scatter, s);
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("permute launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

// Whole-function path (m <= 3, everything of one function in shared memory): one launch per transform.
// Slab tables for k_slab: runs of consecutive planes (= fibres of coordinate 1) with at most
// kSlabThreads lines, kSlabThreads coordinate-1 element-fibres and kSlabElems elements.
int build_slabs(lpfnt_plan *p, const HostIndex &hx) {
    if (p->m < 2 || hx.max_line > kMaxRegT || hx.dims[1].max_lines > kMaxRegT) return LPFNT_OK;
    const DimTable &D = hx.dims[1];
    for (int32_t f = 0; f < D.num_fibres; ++f)
        for (int32_t r = D.fib_ptr[f]; r < D.fib_ptr[f + 1]; ++r)
            if (D.ent[r].off != hx.line_off[r]) return LPFNT_OK;  // planes are not runs of consecutive lines
    std::vector<SlabDesc> desc;
    std::vector<int2> item0(hx.N1), item1;
    item1.reserve(size_t(D.num_threads));
    int max_t = 0;
    int32_t f = 0;
    while (f < D.num_fibres) {
        int32_t f1 = f, nl = 0, nc = 0, ne = 0;
        while (f1 < D.num_fibres) {
            const int32_t pl = D.fib_ptr[f1 + 1] - D.fib_ptr[f1];
            const int32_t pc = D.thr_start[f1 + 1] - D.thr_start[f1];
            const int32_t pe = hx.line_off[D.fib_ptr[f1 + 1]] - hx.line_off[D.fib_ptr[f1]];
            if (f1 > f && (nl + pl > kSlabThreads || nc + pc > kSlabThreads || ne + pe > kSlabElems)) break;
            nl += pl; nc += pc; ne += pe; ++f1;
        }
        if (nl > kSlabThreads || nc > kSlabThreads || ne > kSlabElems) return LPFNT_OK;  // one plane alone is too big
        SlabDesc d{};
        d.l0 = D.fib_ptr[f]; d.nl = nl;
        d.e0 = hx.line_off[d.l0]; d.ne = ne;
        d.c0 = int32_t(item1.size()); d.nc = nc;
        // coordinate 0: the slab's lines, longest first
        std::vector<std::pair<int, int2>> tmp;
        for (int32_t l = d.l0; l < d.l0 + nl; ++l) {
            const int t = hx.line_off[l + 1] - hx.line_off[l];
            tmp.push_back({t, make_int2(t << 24, hx.line_off[l] - d.e0)});
            max_t = std::max(max_t, t);
        }
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, int2> &x, const std::pair<int, int2> &y) { return x.first > y.first; });
        for (int32_t q = 0; q < nl; ++q) item0[d.l0 + q] = tmp[q].second;
        // coordinate 1: (plane, lane) element-fibres, longest first
        tmp.clear();
        for (int32_t g = f; g < f1; ++g) {
            const int32_t p0 = D.fib_ptr[g], pl = D.fib_ptr[g + 1] - p0;
            const int32_t lanes = D.thr_start[g + 1] - D.thr_start[g];
            for (int32_t lane = 0; lane < lanes; ++lane) {
                int t = 0;
                while (t < pl && D.ent[p0 + t].len > lane) ++t;
                tmp.push_back({t, make_int2((32 + (p0 - d.l0)) | (t << 24), lane)});
                max_t = std::max(max_t, t);
            }
        }
        // length classes of 8, natural (plane, lane) order inside a class: a warp's lanes stay on few lines (coalesced
        // result stores, fewer bank conflicts) at ~10 % less uniform lengths -- same trade as for k_whole2
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, int2> &x, const std::pair<int, int2> &y) { return (x.first + 7) / 8 > (y.first + 7) / 8; });
        for (auto &e : tmp) item1.push_back(e.second);
        desc.push_back(d);
        f = f1;
    }
    int rc;
    lpfnt_plan::Slabs &S = p->slabs;
    if ((rc = upload(&S.d_desc, desc.data(), desc.size(), p))) return rc;
    if ((rc = upload(&S.d_item0, item0.data(), item0.size(), p))) return rc;
    if ((rc = upload(&S.d_item1, item1.data(), item1.size(), p))) return rc;
    S.n_slabs = int32_t(desc.size()); S.max_t = max_t;
    // worth it only when the slabs keep the CTA busy (a 2-D function is ONE small plane: 31 lines at n = 30)
    S.ok = hx.N1 / int64_t(desc.size()) >= kSlabThreads / 2;
    return LPFNT_OK;
}

// One pass for coordinates 0 and 1 (k_slab): src -> dst.
int launch_slab_pass(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s) {
    const int mode = sweep_mode(spec);
    const int cap = pick_cap(std::max(p->slabs.max_t, 1));
    std::vector<double> tri;
    repack(spec.packed, p->n + 1, spec.kind, cap, tri);
    SlabArgs a{};
    a.slab = p->slabs.d_desc; a.item0 = p->slabs.d_item0; a.item1 = p->slabs.d_item1; a.line_off = p->d_line_off;
    a.perm_in = gather ? p->d_perm : nullptr;
    a.perm_out = scatter ? p->d_perm : nullptr;
    a.stride = p->N; a.n_slabs = p->slabs.n_slabs;
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 24)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 24, batch - b0));
        a.in = src + b0 * p->N; a.out = dst + b0 * p->N; a.batch = nb;
        const int64_t units = int64_t(a.n_slabs) * nb;
        const int grid = int(std::min<int64_t>(units, int64_t(p->num_sms) * 3));
        cudaError_t e = cudaSuccess;
        trace_begin(p, 1601, s);
        switch (cap) {
            case 8: e = launch_slab<8>(mode, spec.fma, a, tri.data(), grid, s); break;
            case 16: e = launch_slab<16>(mode, spec.fma, a, tri.data(), grid, s); break;
            case 24: e = launch_slab<24>(mode, spec.fma, a, tri.data(), grid, s); break;
            default: e = launch_slab<32>(mode, spec.fma, a, tri.data(), grid, s); break;
        }
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("slab kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

// Box layout for k_whole3, built from A alone (any downward-closed set with m = 2 or 3).
int build_box(lpfnt_plan *p, const int64_t *A, const int64_t *perm, const HostIndex &hx) {
    const int m = p->m;
    const int64_t N = p->N, N1 = p->N1;
    if (m < 2 || m > 3 || N >= 65536 || N1 > kW3Compute || p->n + 1 > kMaxRegT) return LPFNT_OK;
    // (a0, a1, k): m = 3 -> alpha; m = 2 -> (alpha_0, 0, alpha_1)
    auto a1_of = [&](int64_t i) { return m == 3 ? int(A[i * m + 1]) : 0; };
    auto k_of = [&](int64_t i) { return int(A[i * m + m - 1]); };
    int K = 0;
    for (int64_t i = 0; i < N; ++i) K = std::max(K, k_of(i) + 1);
    if (K > 32) return LPFNT_OK;
    std::vector<int> nl(K, 0), len0(K, 0);
    for (int64_t i = 0; i < N; ++i) {
        const int k = k_of(i);
        nl[k] = std::max(nl[k], a1_of(i) + 1);
        len0[k] = std::max(len0[k], int(A[i * m]) + 1);
    }
    lpfnt_plan::Box &B = p->box;
    int64_t W = 0, L = 0;
    for (int k = 0; k < 32; ++k) { B.c.P[k] = 0; B.c.S[k] = 0; B.c.L[k] = 0; }
    for (int k = 0; k < K; ++k) {
        B.c.P[k] = int32_t(W); B.c.S[k] = len0[k] | 1; B.c.L[k] = int32_t(L);
        W += int64_t(nl[k]) * B.c.S[k];
        L += nl[k];
    }
    if (L != N1 || W > 65535 || whole3_smem_bytes(int(W), int(N1), int(N), perm != nullptr) > size_t(227) * 1024) return LPFNT_OK;
    std::vector<uint16_t> lo16(N1);
    std::vector<int2> pline(N1);
    // line lengths per (a1, k)
    std::vector<std::vector<int>> len(K);
    for (int k = 0; k < K; ++k) len[k].assign(nl[k], 0);
    for (int64_t l = 0; l < N1; ++l) {
        const int64_t i = hx.line_off[l];
        lo16[l] = uint16_t(i);
        if (B.c.L[k_of(i)] + a1_of(i) != l) return LPFNT_OK;  // lines are not plane-major (cannot happen for colex order)
        len[k_of(i)][a1_of(i)] = int(hx.line_off[l + 1] - hx.line_off[l]);
        pline[l] = make_int2(int(B.c.P[k_of(i)] + a1_of(i) * B.c.S[k_of(i)]) | (len[k_of(i)][a1_of(i)] << 16), int(i));
    }
    auto sort_desc = [](std::vector<std::pair<int, uint32_t>> &v, std::vector<uint32_t> &out) {
        std::stable_sort(v.begin(), v.end(), [](const std::pair<int, uint32_t> &x, const std::pair<int, uint32_t> &y) { return x.first > y.first; });
        out.resize(v.size());
        for (size_t q = 0; q < v.size(); ++q) out[q] = v[q].second;
    };
    std::vector<std::pair<int, uint32_t>> tmp;
    std::vector<uint32_t> item[3];
    int max_t = 0;
    // coordinate 0: lines
    for (int k = 0; k < K; ++k)
        for (int a1 = 0; a1 < nl[k]; ++a1) {
            const int t = len[k][a1];
            tmp.push_back({t, uint32_t(B.c.P[k] + a1 * B.c.S[k]) | (uint32_t(t) << 16)});
            max_t = std::max(max_t, t);
        }
    sort_desc(tmp, item[0]);
    // coordinate 1 (m = 3): (a0, k)
    tmp.clear();
    if (m == 3)
        for (int k = 0; k < K; ++k)
            for (int a0 = 0; a0 < len0[k]; ++a0) {
                int t = 0;
                while (t < nl[k] && len[k][t] > a0) ++t;
                tmp.push_back({t, uint32_t(B.c.P[k] + a0) | (uint32_t(B.c.S[k]) << 16) | (uint32_t(t) << 24)});
                max_t = std::max(max_t, t);
            }
    sort_desc(tmp, item[1]);
    // last coordinate: (a0, a1) of plane 0
    tmp.clear();
    for (int a1 = 0; a1 < nl[0]; ++a1)
        for (int a0 = 0; a0 < len[0][a1]; ++a0) {
            int t = 0;
            while (t < K && a1 < nl[t] && len[t][a1] > a0) ++t;
            tmp.push_back({t, uint32_t(a0) | (uint32_t(a1) << 8) | (uint32_t(t) << 16)});
            max_t = std::max(max_t, t);
        }
    sort_desc(tmp, item[2]);
    for (int h = 0; h < 3; ++h) if (item[h].size() > size_t(kW3Compute)) return LPFNT_OK;
    int rc;
    for (int h = 0; h < 3; ++h) {
        B.n_item[h] = int32_t(item[h].size());
        if ((rc = upload(&B.d_item[h], item[h].data(), item[h].size(), p))) return rc;
    }
    if ((rc = upload(&B.d_pline, pline.data(), pline.size(), p))) return rc;
    if ((rc = upload(&B.d_line_off16, lo16.data(), lo16.size(), p))) return rc;
    if (perm) {
        std::vector<uint16_t> p16(N);
        for (int64_t i = 0; i < N; ++i) p16[i] = uint16_t(perm[i]);
        if ((rc = upload(&B.d_perm16, p16.data(), p16.size(), p))) return rc;
    }
    B.n_box = int32_t(W); B.max_t = max_t; B.ok = true;
    return LPFNT_OK;
}

bool whole3_ok(const lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims) {
    if (p->use_whole != 3 || p->force_generic || !p->box.ok || int(dims.size()) != p->m) return false;
    for (int q = 0; q < p->m; ++q) if (dims[q] != q) return false;
    if (spec.solve && spec.kind == LPFNT_UPPER) return false;  // fold order changes with the coordinate
    return pick_cap(std::max(p->box.max_t, 1)) > 0;
}

int launch_whole3_fn(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s) {
    const int mode = sweep_mode(spec);
    const int cap = pick_cap(std::max(p->box.max_t, 1));
    std::vector<double> tri;
    repack(spec.packed, p->n + 1, spec.kind, cap, tri);
    const lpfnt_plan::Box &B = p->box;
    Whole3Args a{};
    a.item0 = B.d_item[0]; a.item1 = B.d_item[1]; a.item2 = B.d_item[2];
    a.pline = B.d_pline; a.perm16 = B.d_perm16; a.line_off16 = B.d_line_off16;
    a.gather = gather ? 1 : 0; a.scatter = scatter ? 1 : 0;
    a.n_elem = int32_t(p->N); a.n_box = B.n_box; a.n_lines = int32_t(p->N1); a.m = p->m;
    a.n_item0 = B.n_item[0]; a.n_item1 = B.n_item[1]; a.n_item2 = B.n_item[2];
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 30)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 30, batch - b0));
        a.in = src + b0 * p->N; a.out = dst + b0 * p->N; a.batch = nb;
        const int grid = int(std::min<int64_t>(nb, p->num_sms));
        cudaError_t e = cudaSuccess;
        trace_begin(p, 1500, s);
        switch (cap) {
            case 8: e = launch_whole3<8>(mode, spec.fma, a, B.c, tri.data(), grid, s); break;
            case 16: e = launch_whole3<16>(mode, spec.fma, a, B.c, tri.data(), grid, s); break;
            case 24: e = launch_whole3<24>(mode, spec.fma, a, B.c, tri.data(), grid, s); break;
            default: e = launch_whole3<32>(mode, spec.fma, a, B.c, tri.data(), grid, s); break;
        }
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("whole-function (box) kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

// second-generation whole-function kernel: every sweep in one round, bulk copies need 16-byte granules
bool whole2_ok(const lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, int64_t batch) {
    // automatic: enough items to fill the CTA and enough functions to fill a few SMs -- a single function would run on ONE
    // SM (measured at d=3, n=30, B=1: 29 us against 24 us for the per-coordinate kernels spread over the GPU)
    const bool want = p->use_whole == 2 || (p->use_whole == -1 && p->N1 >= 512 && batch >= 8);
    if (!want || p->force_generic || p->m < 2 || p->m > 3 || int(dims.size()) != p->m) return false;
    for (int q = 0; q < p->m; ++q) if (dims[q] != q) return false;
    if (spec.solve && spec.kind == LPFNT_UPPER) return false;  // fold order changes with the coordinate
    if (p->N >= 65536 || p->N1 > kWhole2Threads || p->n + 1 > kMaxRegT || !p->d_sorted0) return false;
    if (reinterpret_cast<uintptr_t>(in) & 15) return false;
    for (int h = 1; h < p->m; ++h) if (!p->dd[h].sorted || !p->dd[h].ent_off16 || p->dd[h].num_sorted != p->N1) return false;
    return whole2_smem_bytes(int(p->N), int(p->N1), p->has_perm) <= size_t(227) * 1024;
}

bool whole_ok(const lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims) {
    if (p->use_whole != 1 || p->force_generic || p->m > 3 || int(dims.size()) != p->m) return false;
    for (int q = 0; q < p->m; ++q) if (dims[q] != q) return false;
    if (spec.kind != LPFNT_LOWER) return false;
    if (p->N >= 65536 || p->n + 1 > kMaxRegT || !p->d_sorted0) return false;
    for (int h = 1; h < p->m; ++h) if (!p->dd[h].sorted || !p->dd[h].ent_off16) return false;
    return whole_smem_bytes(int(p->N), int(p->N1)) <= size_t(220) * 1024;
}

int launch_whole_fn(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s, bool gen2 = false) {
    const int n1 = p->n + 1;
    const int mode = sweep_mode(spec);
    int len = p->max_line;
    for (int h = 1; h < p->m; ++h) len = std::max(len, p->dd[h].max_lines);
    const int cap = pick_cap(std::max(len, 1));
    std::vector<double> tri;
    repack(spec.packed, n1, spec.kind, cap, tri);
    WholeArgs a{};
    a.in = src; a.out = dst;
    a.sorted[0] = p->d_sorted0; a.ent_off[0] = nullptr;
    for (int h = 1; h < p->m; ++h) { a.sorted[h] = p->dd[h].sorted; a.ent_off[h] = p->dd[h].ent_off16; }
    if (gen2) {
        const int ord = spec.fma ? p->w2_order_fma : p->w2_order_exact;
        for (int h = 0; h < p->m; ++h)
            if (p->w2_items[h][(ord >> (2 * h)) & 3]) a.sorted[h] = p->w2_items[h][(ord >> (2 * h)) & 3];
    }
    a.perm_in = gather ? p->d_perm : nullptr;
    a.perm_out = scatter ? p->d_perm : nullptr;
    a.n_elem = int(p->N); a.n_lines = int(p->N1); a.m = p->m;
    a.dbg = gen2 ? p->d_timeline : nullptr;
    const size_t smem = whole_smem_bytes(int(p->N), int(p->N1));
    const int threads = spec.fma ? 256 : 384;
    int occ = int(std::min<size_t>((size_t(227) * 1024) / (smem + 1024), size_t(2048 / threads)));
    occ = std::max(1, std::min(occ, spec.fma ? 2 : 1 + (smem < 100 * 1024)));
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 30)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 30, batch - b0));
        a.in = src + b0 * p->N; a.out = dst + b0 * p->N; a.batch = nb;
        const int grid = int(std::min<int64_t>(nb, int64_t(p->num_sms) * (gen2 ? 1 : occ)));
        cudaError_t e = cudaSuccess;
        trace_begin(p, gen2 ? 1400 : 1000, s);
        if (gen2) {
            switch (cap) {
                case 8: e = launch_whole2<8>(mode, spec.fma, a, tri.data(), grid, s); break;
                case 16: e = launch_whole2<16>(mode, spec.fma, a, tri.data(), grid, s); break;
                case 24: e = launch_whole2<24>(mode, spec.fma, a, tri.data(), grid, s); break;
                default: e = launch_whole2<32>(mode, spec.fma, a, tri.data(), grid, s); break;
            }
        } else
        switch (cap) {
            case 8: e = launch_whole<8>(mode, spec.fma, a, tri.data(), grid, s); break;
            case 16: e = launch_whole<16>(mode, spec.fma, a, tri.data(), grid, s); break;
            case 24: e = launch_whole<24>(mode, spec.fma, a, tri.data(), grid, s); break;
            default: e = launch_whole<32>(mode, spec.fma, a, tri.data(), grid, s); break;
        }
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("whole-function kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

// Device-resident pipeline: sweeps along `dims` (ascending), optional colex gather on the way
// in and scatter on the way out.  in/out may alias.
//
// Pass plan: with the tile kernels the coordinate-0 sweep rides along with the coordinate-1 pass
// (one pass fewer); every remaining coordinate is one pass.  Passes 2.. run in place.  A batched
// multi-pass transform is cut into chunks of functions of ~l2_chunk_bytes so that the passes
// after the first find their input in the 126 MB L2 instead of HBM.
struct Pass { int h; bool fuse0; };

int run_device(lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, double *out,
               int64_t batch, bool gather, bool scatter, cudaStream_t s, int mid_slot = 2) {
    gather = gather && p->has_perm;
    scatter = scatter && p->has_perm;
    if (dims.empty()) { set_error("no sweep dimensions"); return LPFNT_EINVAL; }
    if (spec.solve && spec.kind == LPFNT_UPPER && (p->use_tiles || p->use_sorted || p->use_planes || p->use_pipe)) {
        set_error("the upper triangular solve only runs on the default line/group kernels");
        return LPFNT_EUNSUPPORTED;
    }
    if (gather && dims[0] != 0) { set_error("gather requires a sweep along coordinate 0"); return LPFNT_EINVAL; }
    if (whole3_ok(p, spec, dims)) {
        // in-place is safe: a function is complete in shared memory / registers before its first store
        return launch_whole3_fn(p, spec, in, out, batch, gather, scatter, s);
    }
    if (whole2_ok(p, spec, dims, in, batch)) {
        // in-place is safe: a function is complete in shared memory / registers before its first store
        return launch_whole_fn(p, spec, in, out, batch, gather, scatter, s, true);
    }
    if (whole_ok(p, spec, dims)) {
        // in-place is safe: a function is read completely into shared memory before it is written
        return launch_whole_fn(p, spec, in, out, batch, gather, scatter, s);
    }
    std::vector<Pass> passes;
    {
        size_t q = 0;
        const bool can_plane = !p->force_generic && p->use_planes && p->num_planes > 0 && pick_cap(std::max(p->max_line, p->dd[1].max_lines)) > 0;
        const bool can_slab = p->use_slab && p->slabs.ok && !p->force_generic && !p->use_tiles && !p->use_planes && !p->use_sorted && p->m >= 2 &&
                              pick_cap(std::max(p->slabs.max_t, 1)) > 0 && !(spec.solve && spec.kind == LPFNT_UPPER);
        const bool can_tile = can_slab || can_plane || !p->force_generic && !p->use_sorted && p->use_tiles && p->m >= 2 && p->dd[1].num_tiles > 0 &&
                              pick_cap(std::max(p->max_line, p->dd[1].max_lines)) > 0;
        if (dims.size() >= 2 && dims[0] == 0 && dims[1] == 1 && can_tile) { passes.push_back({1, true}); q = 2; }
        for (; q < dims.size(); ++q) passes.push_back({dims[q], false});
    }
    const int np = int(passes.size());
    const size_t fbytes = size_t(p->N) * sizeof(double);
    int64_t chunk = batch;
    const bool permuting = gather || scatter;
    bool pipe = !p->force_generic && p->use_tiles && p->use_pipe && !p->use_sorted;
    for (const Pass &ps : passes)
        pipe = pipe && ps.h >= 1 && p->dd[ps.h].num_tiles > 0 && pick_cap(std::max(p->dd[ps.h].max_lines, ps.fuse0 ? p->max_line : 0)) > 0;
    const int nlaunch = np + (pipe ? int(gather) + int(scatter) : 0);
    if (nlaunch >= 2 && p->l2_chunk_bytes > 0) chunk = std::max<int64_t>(1, std::min<int64_t>(batch, p->l2_chunk_bytes / int64_t(fbytes)));

    if (pipe) {
        // stand-alone permutation passes + pipelined sweeps; everything after the first launch of a
        // chunk finds its input in L2
        const bool need_mid = permuting;  // the permutation passes never run in place
        double *mid_base = nullptr;
        if (need_mid) { int rc = ensure_ws(p, mid_slot, size_t(chunk) * fbytes); if (rc) return rc; mid_base = p->ws[mid_slot]; }
        for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
            const int64_t nb = std::min(chunk, batch - b0);
            const double *cur = in + b0 * p->N;
            double *cout = out + b0 * p->N;
            if (gather) { int rc = launch_perm(p, cur, mid_base, nb, false, s); if (rc) return rc; cur = mid_base; }
            for (int q = 0; q < np; ++q) {
                double *dst = (q == np - 1 && !scatter) ? cout : (need_mid ? mid_base : cout);
                int rc = launch_pipe(p, spec, passes[q].h, passes[q].fuse0, cur, dst, nb, s);
                if (rc) return rc;
                cur = dst;
            }
            if (scatter) { int rc = launch_perm(p, cur, cout, nb, true, s); if (rc) return rc; }
        }
        return LPFNT_OK;
    }

    // a permuting pass must not run in place; with >= 2 passes the middle buffer absorbs that
    const bool need_mid = (np == 1) ? (permuting && in == out) : (scatter || (permuting && in == out));
    double *mid_base = nullptr;
    if (need_mid) { int rc = ensure_ws(p, mid_slot, size_t(chunk) * fbytes); if (rc) return rc; mid_base = p->ws[mid_slot]; }
    for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
        const int64_t nb = std::min(chunk, batch - b0);
        const double *cin = in + b0 * p->N;
        double *cout = out + b0 * p->N;
        if (np == 1) {
            double *dst = need_mid ? mid_base : cout;
            int rc = launch_dim(p, spec, passes[0].h, passes[0].fuse0, cin, dst, nb, gather, scatter, s);
            if (rc) return rc;
            if (dst != cout) CK(cudaMemcpyAsync(cout, dst, size_t(nb) * fbytes, cudaMemcpyDeviceToDevice, s));
            continue;
        }
        double *mid = need_mid ? mid_base : cout;
        for (int q = 0; q < np; ++q) {
            const bool first = (q == 0), last = (q == np - 1);
            int rc = launch_dim(p, spec, passes[q].h, passes[q].fuse0, first ? cin : mid, last ? cout : mid, nb,
                                first && gather, last && scatter, s);
            if (rc) return rc;
        }
    }
    return LPFNT_OK;
}

// Host-pointer front end: stage through ws[0]/ws[1] in chunks of functions.
// Sequence of sweeps (one spec = one triangular operator applied along `dims`); gather rides on the first,
// scatter on the last.
int run_specs(lpfnt_plan *p, const std::vector<SweepSpec> &specs, const std::vector<int> &dims, const double *in, double *out,
              int64_t batch, bool gather, bool scatter, cudaStream_t s, int mid_slot) {
    const int ns = int(specs.size());
    for (int i = 0; i < ns; ++i) {
        int rc = run_device(p, specs[i], dims, i == 0 ? in : out, out, batch, gather && i == 0, scatter && i == ns - 1, s, mid_slot);
        if (rc) return rc;
    }
    return LPFNT_OK;
}

int run_any(lpfnt_plan *p, const std::vector<SweepSpec> &specs, const std::vector<int> &dims, const double *in, double *out,
            int64_t batch, bool gather, bool scatter, int on_device, void *stream) {
    if (!p || !in || !out || batch < 1) { set_error("null pointer or empty batch"); return LPFNT_EINVAL; }
    for (const SweepSpec &sp : specs)
        if (!sp.packed) { set_error("the plan was created without the matrix this operation needs"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(p->device));
    if (on_device) return run_specs(p, specs, dims, in, out, batch, gather, scatter, static_cast<cudaStream_t>(stream), 2);
    // Host pointers: cut the batch into chunks and run them on two lanes (streams) with their own staging
    // buffers, so that the H2D copy of chunk i+1 overlaps the sweeps and the D2H copy of chunk i
    // (PCIe is full duplex; the copies dominate this path by an order of magnitude).
    const size_t fbytes = size_t(p->N) * sizeof(double);
    // measured on B200 / PCIe Gen5 (tools/e2e.py): 16 MB chunks are best for a 125 MB batch (5.0 vs 4.3 GDOF/s with
    // 64 MB), 32 MB for 500 MB; below 8 MB the per-chunk launch and copy overheads take over
    const size_t cbytes = p->host_chunk_bytes > 0 ? size_t(p->host_chunk_bytes)
                                                  : std::min<size_t>(size_t(32) << 20, std::max<size_t>(size_t(16) << 20, size_t(batch) * fbytes / 8));
    int64_t chunk = std::max<int64_t>(1, int64_t(cbytes / fbytes));
    if (batch < 2 * chunk) chunk = std::max<int64_t>(1, (batch + 1) / 2);
    const int lanes = batch > chunk ? 2 : 1;
    for (int l = 0; l < lanes; ++l) {
        int rc = ensure_ws(p, 3 * l, size_t(chunk) * fbytes);
        if (rc) return rc;
        rc = ensure_ws(p, 3 * l + 1, size_t(chunk) * fbytes);
        if (rc) return rc;
    }
    int i = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += chunk, ++i) {
        const int64_t nb = std::min(chunk, batch - b0);
        const int l = (lanes == 2) ? (i & 1) : 0;
        cudaStream_t s = l ? p->stream2 : p->stream;
        CK(cudaMemcpyAsync(p->ws[3 * l], in + b0 * p->N, size_t(nb) * fbytes, cudaMemcpyHostToDevice, s));
        int rc = run_specs(p, specs, dims, p->ws[3 * l], p->ws[3 * l + 1], nb, gather, scatter, s, 3 * l + 2);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out + b0 * p->N, p->ws[3 * l + 1], size_t(nb) * fbytes, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(p->stream));
    if (lanes == 2) CK(cudaStreamSynchronize(p->stream2));
    return LPFNT_OK;
}

int run_any(lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, double *out,
            int64_t batch, bool gather, bool scatter, int on_device, void *stream) {
    return run_any(p, std::vector<SweepSpec>{spec}, dims, in, out, batch, gather, scatter, on_device, stream);
}

std::vector<int> all_dims(int m) {
    std::vector<int> d(m);
    for (int i = 0; i < m; ++i) d[i] = i;
    return d;
}

}  // namespace

extern "C" {

const char *lpfnt_version(void) { return "lpfnt-b200 0.1.0 (sm_100a)"; }

int lpfnt_device_count(int *count) {
    if (!count) { set_error("null pointer"); return LPFNT_EINVAL; }
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    return LPFNT_OK;
}

int lpfnt_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) { set_error("bad arguments"); return LPFNT_EINVAL; }
    CK(cudaHostAlloc(ptr, size_t(std::max<int64_t>(bytes, 1)), cudaHostAllocPortable));
    return LPFNT_OK;
}
int lpfnt_host_free(void *ptr) {
    if (ptr) CK(cudaFreeHost(ptr));
    return LPFNT_OK;
}

int lpfnt_plan_create(int m, int n, int64_t N, const int64_t *A, const int64_t *perm, const double *nodes,
                      const double *Vx_lt, const double *inv_Vx_lt, const double *const *Dx_ut,
                      const double *const *DxT_lt, int device, lpfnt_plan **plan) {
    if (!plan) { set_error("plan pointer is null"); return LPFNT_EINVAL; }
    *plan = nullptr;
    if (!A || m < 1 || n < 0 || N < 1) { set_error("plan_create: need A, m >= 1, n >= 0, N >= 1"); return LPFNT_EINVAL; }
    if (m > kMaxDim) { set_error("plan_create: spatial dimension above the supported maximum (12)"); return LPFNT_EUNSUPPORTED; }
    HostIndex hx;
    if (!hx.build(A, N, m, n)) return LPFNT_EINVAL;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("plan_create: no such CUDA device"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(device));

    lpfnt_plan *p = new lpfnt_plan();
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) p->num_sms = prop.multiProcessorCount;
    }
    p->device = device; p->m = m; p->n = n; p->N = N; p->N1 = hx.N1; p->max_line = hx.max_line;
    const int ts = tri_size(n + 1);
    if (nodes) p->nodes.assign(nodes, nodes + n + 1);
    if (Vx_lt) p->Vx_lt.assign(Vx_lt, Vx_lt + ts);
    if (inv_Vx_lt) p->inv_Vx_lt.assign(inv_Vx_lt, inv_Vx_lt + ts);
    for (int k = 0; k < 3; ++k) {
        if (Dx_ut && Dx_ut[k]) p->Dx_ut[k].assign(Dx_ut[k], Dx_ut[k] + ts);
        if (DxT_lt && DxT_lt[k]) p->DxT_lt[k].assign(DxT_lt[k], DxT_lt[k] + ts);
    }
    int rc = LPFNT_OK;
    auto fail = [&](int code) { lpfnt_plan_destroy(p); return code; };
    if ((rc = upload(&p->d_line_off, hx.line_off.data(), hx.line_off.size(), p))) return fail(rc);
    if (perm) {
        std::vector<int32_t> p32(N);
        std::vector<uint8_t> seen(N, 0);
        for (int64_t i = 0; i < N; ++i) {
            if (perm[i] < 0 || perm[i] >= N || seen[perm[i]]) { set_error("plan_create: perm is not a permutation of 0..N-1"); return fail(LPFNT_EINVAL); }
            seen[perm[i]] = 1;
            p32[i] = int32_t(perm[i]);
        }
        if ((rc = upload(&p->d_perm, p32.data(), p32.size(), p))) return fail(rc);
        p->has_perm = true;
    }
    for (int h = 1; h < m; ++h) {
        const DimTable &D = hx.dims[h];
        DeviceDim &d = p->dd[h];
        d.num_threads = int32_t(D.num_threads);
        d.max_lines = D.max_lines;
        if ((rc = upload(&d.thr_fib, D.thr_fib.data(), D.thr_fib.size(), p))) return fail(rc);
        if ((rc = upload(&d.thr_start, D.thr_start.data(), D.thr_start.size(), p))) return fail(rc);
        if ((rc = upload(&d.fib_ptr, D.fib_ptr.data(), D.fib_ptr.size(), p))) return fail(rc);
        static_assert(sizeof(FibEnt) == sizeof(int2), "FibEnt must match int2");
        if ((rc = upload(&d.ent, reinterpret_cast<const int2 *>(D.ent.data()), D.ent.size(), p))) return fail(rc);
        static_assert(sizeof(SortedItem) == sizeof(int2), "SortedItem must match int2");
        if (!D.sorted.empty()) {
            d.num_sorted = int32_t(D.sorted.size());
            if ((rc = upload(&d.sorted, reinterpret_cast<const int2 *>(D.sorted.data()), D.sorted.size(), p))) return fail(rc);
        }
        if (N < 65536) {
            std::vector<uint16_t> eo(D.ent.size());
            for (size_t r = 0; r < D.ent.size(); ++r) eo[r] = uint16_t(D.ent[r].off);
            if ((rc = upload(&d.ent_off16, eo.data(), eo.size(), p))) return fail(rc);
        }
        if (!D.groups.empty()) {
            d.num_groups = int32_t(D.groups.size());
            if ((rc = upload(&d.groups, reinterpret_cast<const int2 *>(D.groups.data()), D.groups.size(), p))) return fail(rc);
            if ((rc = upload(&d.group_t, D.group_t.data(), D.group_t.size(), p))) return fail(rc);
            if ((rc = upload(&d.group_cta_cap, D.group_cta_cap.data(), D.group_cta_cap.size(), p))) return fail(rc);
        }
        if (!D.tile_fib.empty()) {
            d.num_tiles = int32_t(D.tile_fib.size()) - 1;
            if ((rc = upload(&d.tile_fib, D.tile_fib.data(), D.tile_fib.size(), p))) return fail(rc);
            if ((rc = upload(&d.items, D.items.data(), D.items.size(), p))) return fail(rc);
            if ((rc = upload(&d.rows0, D.rows0.data(), D.rows0.size(), p))) return fail(rc);
            std::vector<int4> rec(d.num_tiles);
            for (int tl = 0; tl < d.num_tiles; ++tl) {
                const int f0 = D.tile_fib[tl], f1 = D.tile_fib[tl + 1];
                rec[tl] = make_int4(D.fib_ptr[f0], D.fib_ptr[f1] - D.fib_ptr[f0], D.thr_start[f0], D.thr_start[f1] - D.thr_start[f0]);
            }
            if ((rc = upload(&d.tile_rec, rec.data(), rec.size(), p))) return fail(rc);
            {
                std::vector<int2> te(d.num_tiles);
                for (int tl = 0; tl < d.num_tiles; ++tl) te[tl] = make_int2(D.tile_elem[tl], D.tile_nelem[tl]);
                if ((rc = upload(&d.tile_elem, te.data(), te.size(), p))) return fail(rc);
            }
            if ((rc = upload(&d.row_pos, D.row_pos.data(), D.row_pos.size(), p))) return fail(rc);
            {
                std::vector<int2> rd(D.ent.size());
                for (size_t r = 0; r < D.ent.size(); ++r) rd[r] = make_int2(D.ent[r].off - int(D.row_pos[r]), int(D.row_pos[r]));
                if ((rc = upload(&d.row_desc, rd.data(), rd.size(), p))) return fail(rc);
            }
            if ((rc = upload(&d.elem_row, D.elem_row.data(), D.elem_row.size(), p))) return fail(rc);
        }
    }
    if ((rc = upload(&p->d_line_order, hx.line_order.data(), hx.line_order.size(), p))) return fail(rc);
    if (m >= 2 && hx.dims[1].max_lines <= 32 && hx.max_line <= 32) {
        // the fibres of coordinate 1 are the planes; their rows are consecutive lines, so the CSR of the
        // fibres is also the CSR of lines per plane
        const DimTable &D1 = hx.dims[1];
        bool consecutive = true;
        for (int32_t f = 0; f < D1.num_fibres && consecutive; ++f)
            for (int32_t r = D1.fib_ptr[f]; r < D1.fib_ptr[f + 1]; ++r)
                if (D1.ent[r].off != hx.line_off[r]) { consecutive = false; break; }
        if (consecutive) {
            if ((rc = upload(&p->d_plane_line, D1.fib_ptr.data(), D1.fib_ptr.size(), p))) return fail(rc);
            p->num_planes = D1.num_fibres;
        }
    }
    if (m >= 2 && m <= 3 && N < 65536 && hx.max_line <= 255) {
        for (int h = 0; h < m; ++h) {
            std::vector<std::pair<int, int2>> nat;  // natural order with heights
            if (h == 0) {
                for (int64_t l = 0; l < hx.N1; ++l) {
                    const int t = hx.line_off[l + 1] - hx.line_off[l];
                    nat.push_back({t, make_int2(hx.line_off[l], t << 8)});
                }
            } else {
                const DimTable &D = hx.dims[h];
                if (D.max_lines > 255) continue;
                for (int32_t f = 0; f < D.num_fibres; ++f) {
                    const int32_t p0 = D.fib_ptr[f], nl = D.fib_ptr[f + 1] - p0;
                    const int32_t lanes = D.thr_start[f + 1] - D.thr_start[f];
                    for (int32_t lane = 0; lane < lanes; ++lane) {
                        int32_t t = 0;
                        while (t < nl && D.ent[p0 + t].len > lane) ++t;
                        nat.push_back({t, make_int2(p0, lane | (t << 8))});
                    }
                }
            }
            const int cls[4] = {1, 8, 16, 1 << 20};
            for (int v = 0; v < 4; ++v) {
                std::vector<std::pair<int, int2>> s = nat;
                const int q = cls[v];
                std::stable_sort(s.begin(), s.end(), [q](const std::pair<int, int2> &x, const std::pair<int, int2> &y) { return (x.first + q - 1) / q > (y.first + q - 1) / q; });
                std::vector<int2> out(s.size());
                for (size_t i = 0; i < s.size(); ++i) out[i] = s[i].second;
                if ((rc = upload(&p->w2_items[h][v], out.data(), out.size(), p))) return fail(rc);
            }
        }
    }
    if ((rc = build_box(p, A, perm, hx))) return fail(rc);
    if ((rc = build_slabs(p, hx))) return fail(rc);
    if (!hx.sorted0.empty())
        if ((rc = upload(&p->d_sorted0, reinterpret_cast<const int2 *>(hx.sorted0.data()), hx.sorted0.size(), p))) return fail(rc);
    if (!hx.eval_block.empty()) {
        static_assert(sizeof(EvalBlock) == sizeof(int4), "EvalBlock must match int4");
        if ((rc = upload(&p->d_eval_line, reinterpret_cast<const int2 *>(hx.eval_line.data()), hx.eval_line.size(), p))) return fail(rc);
        if ((rc = upload(&p->d_eval_block, reinterpret_cast<const int4 *>(hx.eval_block.data()), hx.eval_block.size(), p))) return fail(rc);
        p->num_eval_blocks = int32_t(hx.eval_block.size());
        if (cudaMalloc(reinterpret_cast<void **>(&p->d_cpad), sizeof(double) * size_t(hx.eval_padded + 4)) != cudaSuccess) { set_error("cudaMalloc(cpad) failed"); return fail(LPFNT_ENOMEM); }
        p->table_bytes += int64_t(sizeof(double)) * (hx.eval_padded + 4);
    }
    if (!hx.line_alpha.empty())
        if ((rc = upload(&p->d_line_alpha, hx.line_alpha.data(), hx.line_alpha.size(), p))) return fail(rc);
    // the generic eval needs A on the device; keep a host copy when that path can be reached
    if (n + 1 > kMaxRegT || m > kEvalMaxM || n > 255 || N * m <= (int64_t(1) << 26)) p->hA.assign(A, A + N * m);
    if (nodes && (rc = upload(&p->d_nodes, nodes, size_t(n + 1), p))) return fail(rc);
    {
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p->d_mat), sizeof(double) * ts);
        if (e != cudaSuccess) { set_error("cudaMalloc(d_mat) failed"); return fail(LPFNT_ENOMEM); }
        p->table_bytes += sizeof(double) * ts;
    }
    {
        cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking);
        if (e != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(LPFNT_ECUDA); }
    }
    *plan = p;
    return LPFNT_OK;
}

int lpfnt_plan_destroy(lpfnt_plan *p) {
    if (!p) return LPFNT_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_eval_line); cudaFree(p->d_eval_block); cudaFree(p->d_cpad);
    for (int h = 0; h < 3; ++h) cudaFree(p->box.d_item[h]);
    for (int h = 0; h < 3; ++h) for (int v = 0; v < 4; ++v) cudaFree(p->w2_items[h][v]);
    cudaFree(p->d_timeline);
    cudaFree(p->slabs.d_desc); cudaFree(p->slabs.d_item0); cudaFree(p->slabs.d_item1);
    cudaFree(p->box.d_pline); cudaFree(p->box.d_perm16); cudaFree(p->box.d_line_off16);
    cudaFree(p->d_line_off); cudaFree(p->d_perm); cudaFree(p->d_sorted0); cudaFree(p->d_plane_line); cudaFree(p->d_line_order); cudaFree(p->d_line_alpha); cudaFree(p->d_A);
    cudaFree(p->d_nodes); cudaFree(p->d_mat);
    for (int h = 0; h <= kMaxDim; ++h) {
        cudaFree(p->dd[h].thr_fib); cudaFree(p->dd[h].thr_start); cudaFree(p->dd[h].fib_ptr); cudaFree(p->dd[h].ent);
        cudaFree(p->dd[h].tile_fib); cudaFree(p->dd[h].items); cudaFree(p->dd[h].rows0); cudaFree(p->dd[h].tile_rec); cudaFree(p->dd[h].sorted); cudaFree(p->dd[h].ent_off16); cudaFree(p->dd[h].groups); cudaFree(p->dd[h].group_t); cudaFree(p->dd[h].group_cta_cap); cudaFree(p->dd[h].tile_elem); cudaFree(p->dd[h].row_pos); cudaFree(p->dd[h].row_desc); cudaFree(p->dd[h].elem_row);
    }
    for (int s = 0; s < 6; ++s) cudaFree(p->ws[s]);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->stream2) cudaStreamDestroy(p->stream2);
    delete p;
    return LPFNT_OK;
}

int64_t lpfnt_plan_device_bytes(const lpfnt_plan *p) {
    if (!p) return 0;
    int64_t b = p->table_bytes;
    for (int s = 0; s < 6; ++s) b += int64_t(p->ws_bytes[s]);
    return b;
}
int64_t lpfnt_plan_launch_count(const lpfnt_plan *p) { return p ? p->launches : 0; }

int lpfnt_plan_set_option(lpfnt_plan *p, const char *name, int64_t value) {
    if (!p || !name) { set_error("bad arguments"); return LPFNT_EINVAL; }
    if (std::strcmp(name, "force_generic") == 0) { p->force_generic = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "trace") == 0) { p->trace = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "eval_block") == 0) { p->use_eval_block = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "use_planes") == 0) { p->use_planes = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "host_chunk_bytes") == 0) { p->host_chunk_bytes = value <= 0 ? 0 : std::max<int64_t>(value, 1 << 16); return LPFNT_OK; }
    if (std::strcmp(name, "w2_order_exact") == 0) { p->w2_order_exact = int(value); return LPFNT_OK; }
    if (std::strcmp(name, "w2_order_fma") == 0) { p->w2_order_fma = int(value); return LPFNT_OK; }
    if (std::strcmp(name, "use_cta_cap") == 0) { p->use_cta_cap = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "timeline") == 0) {
        if (value && !p->d_timeline) {
            CK(cudaMalloc(reinterpret_cast<void **>(&p->d_timeline), 4 * 24 * 16 * sizeof(long long)));
            CK(cudaMemset(p->d_timeline, 0, 4 * 24 * 16 * sizeof(long long)));
        } else if (!value && p->d_timeline) { cudaFree(p->d_timeline); p->d_timeline = nullptr; }
        return LPFNT_OK;
    }
    if (std::strcmp(name, "use_slab") == 0) {
        p->use_slab = value ? 1 : 0;
        if (value == 2 && p->slabs.n_slabs > 0) p->slabs.ok = true;  // 2: force the slab kernel even for sparsely filled slabs (tests)
        return LPFNT_OK;
    }
    if (std::strcmp(name, "use_whole") == 0) { p->use_whole = value < -1 ? -1 : (value > 3 ? 3 : value); return LPFNT_OK; }
    if (std::strcmp(name, "use_two_lanes") == 0) { p->use_two_lanes = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "use_groups") == 0) { p->use_groups = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "use_sorted") == 0) { p->use_sorted = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "debug_skip") == 0) { p->debug_skip = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "use_tiles") == 0) { p->use_tiles = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "use_pipe") == 0) { p->use_pipe = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "l2_chunk_bytes") == 0) { p->l2_chunk_bytes = value; return LPFNT_OK; }
    set_error(std::string("unknown option: ") + name);
    return LPFNT_EINVAL;
}

int lpfnt_embed_apply(int device, const int64_t *phi, int64_t n_small, int64_t n_big, const double *in, double *out, int64_t batch,
                      int restrict_, int on_device, void *stream) {
    if (!phi || !in || !out || n_small < 1 || n_big < n_small || batch < 1) { set_error("embed_apply: bad arguments"); return LPFNT_EINVAL; }
    for (int64_t e = 0; e < n_small; ++e)
        if (phi[e] < 0 || phi[e] >= n_big) { set_error("embed_apply: phi is not an index map into the fine space"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n_in = restrict_ ? n_big : n_small, n_out = restrict_ ? n_small : n_big;
    int64_t *d_phi = nullptr;
    double *d_in = nullptr, *d_out = nullptr;
    int rc = LPFNT_OK;
    auto cleanup = [&]() { cudaFree(d_phi); if (!on_device) { cudaFree(d_in); cudaFree(d_out); } };
#define CKE(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e__)); cleanup(); return LPFNT_ECUDA; } } while (0)
    CKE(cudaMalloc(reinterpret_cast<void **>(&d_phi), sizeof(int64_t) * size_t(n_small)));
    CKE(cudaMemcpyAsync(d_phi, phi, sizeof(int64_t) * size_t(n_small), cudaMemcpyHostToDevice, s));
    if (on_device) { d_in = const_cast<double *>(in); d_out = out; }
    else {
        CKE(cudaMalloc(reinterpret_cast<void **>(&d_in), sizeof(double) * size_t(n_in * batch)));
        CKE(cudaMalloc(reinterpret_cast<void **>(&d_out), sizeof(double) * size_t(n_out * batch)));
        CKE(cudaMemcpyAsync(d_in, in, sizeof(double) * size_t(n_in * batch), cudaMemcpyHostToDevice, s));
    }
    if (!restrict_) CKE(cudaMemsetAsync(d_out, 0, sizeof(double) * size_t(n_out * batch), s));
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = int(std::min<int64_t>(65535, batch - b0));
        CKE(launch_embed(d_in + b0 * n_in, d_out + b0 * n_out, d_phi, n_small, n_big, nb, restrict_ != 0, s));
    }
    if (!on_device) CKE(cudaMemcpyAsync(out, d_out, sizeof(double) * size_t(n_out * batch), cudaMemcpyDeviceToHost, s));
    CKE(cudaStreamSynchronize(s));  // phi (and the staging buffers) are freed below
#undef CKE
    cleanup();
    return rc;
}

int lpfnt_plan_timeline_read(lpfnt_plan *p, long long *out, int count) {
    if (!p || !out || !p->d_timeline) { set_error("timeline: not enabled"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, p->d_timeline, sizeof(long long) * size_t(std::min(count, 4 * 24 * 16)), cudaMemcpyDeviceToHost));
    return LPFNT_OK;
}

int lpfnt_plan_trace_read(lpfnt_plan *p, int max_records, int *labels, float *ms, int *count) {
    if (!p || !count) { set_error("bad arguments"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(p->device));
    int n = 0;
    for (auto &r : p->trace_recs) {
        if (r.b) {
            CK(cudaEventSynchronize(r.b));
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, r.a, r.b));
            if (n < max_records && labels && ms) { labels[n] = r.label; ms[n] = t; }
            ++n;
        }
        cudaEventDestroy(r.a);
        if (r.b) cudaEventDestroy(r.b);
    }
    p->trace_recs.clear();
    *count = n;
    return LPFNT_OK;
}

int lpfnt_itransform(lpfnt_plan *p, const double *packed, int kind, int arith, const double *in, double *out,
                     int64_t batch, int gather_perm, int scatter_perm, int on_device, void *stream) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (kind != LPFNT_LOWER && kind != LPFNT_UPPER) { set_error("kind must be LPFNT_LOWER or LPFNT_UPPER"); return LPFNT_EINVAL; }
    SweepSpec spec{packed, kind, false, arith == LPFNT_ARITH_FMA};
    return run_any(p, spec, all_dims(p->m), in, out, batch, gather_perm != 0, scatter_perm != 0, on_device, stream);
}

int lpfnt_sweep(lpfnt_plan *p, const double *packed, int kind, int solve, int arith, uint32_t dims_mask, const double *in, double *out,
                int64_t batch, int on_device, void *stream) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (kind != LPFNT_LOWER && kind != LPFNT_UPPER) { set_error("kind must be LPFNT_LOWER or LPFNT_UPPER"); return LPFNT_EINVAL; }
    std::vector<int> dims;
    for (int h = 0; h < p->m; ++h) if (dims_mask >> h & 1u) dims.push_back(h);
    if (dims.empty() || (p->m < 32 && (dims_mask >> p->m) != 0)) { set_error("sweep: dims_mask selects no coordinate or one beyond m - 1"); return LPFNT_EINVAL; }
    SweepSpec spec{packed, kind, solve != 0, solve == 0 && arith == LPFNT_ARITH_FMA};
    return run_any(p, spec, dims, in, out, batch, false, false, on_device, stream);
}

int lpfnt_transform(lpfnt_plan *p, const double *packed, int kind, const double *in, double *out, int64_t batch,
                    int gather_perm, int scatter_perm, int on_device, void *stream) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (kind != LPFNT_LOWER) { set_error("transform: only the lower (Newton) solve is implemented"); return LPFNT_EUNSUPPORTED; }
    SweepSpec spec{packed, kind, true, false};
    return run_any(p, spec, all_dims(p->m), in, out, batch, gather_perm != 0, scatter_perm != 0, on_device, stream);
}

int lpfnt_dtransform(lpfnt_plan *p, const double *packed, int kind, int arith, int i, const double *in, double *out,
                     int64_t batch, int on_device, void *stream) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (i < 0 || i >= p->m) {  // molecules.py:279-280
        set_error("Choose a coordinate between i=0 and i=" + std::to_string(p->m - 1) + ".");
        return LPFNT_EINVAL;
    }
    if (kind != LPFNT_LOWER && kind != LPFNT_UPPER) { set_error("kind must be LPFNT_LOWER or LPFNT_UPPER"); return LPFNT_EINVAL; }
    SweepSpec spec{packed, kind, false, arith == LPFNT_ARITH_FMA};
    return run_any(p, spec, std::vector<int>{i}, in, out, batch, false, false, on_device, stream);
}

int lpfnt_plan_set_upper_factors(lpfnt_plan *p, const double *Vx_ut, const double *inv_Vx_ut, int chebyshev_eval) {
    if (!p || !Vx_ut) { set_error("set_upper_factors: need the packed upper factor of V"); return LPFNT_EINVAL; }
    const int ts = tri_size(p->n + 1);
    p->Vx_ut.assign(Vx_ut, Vx_ut + ts);
    if (inv_Vx_ut) p->inv_Vx_ut.assign(inv_Vx_ut, inv_Vx_ut + ts);
    else p->inv_Vx_ut.clear();
    p->chebyshev_eval = chebyshev_eval ? 1 : 0;
    return LPFNT_OK;
}

int lpfnt_fnt(lpfnt_plan *p, const double *in, double *out, int64_t batch, int on_device, void *stream) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (!p->Vx_ut.empty()) {
        // V = L U (transform.py:491-519): lower mat-vec with inv(L), then upper mat-vec with inv(U)
        if (p->inv_Vx_lt.empty() || p->inv_Vx_ut.empty()) {
            // precomputation=False (transform.py:520-548): forward substitution with L, then backward with U
            std::vector<SweepSpec> specs{{p->Vx_lt.empty() ? nullptr : p->Vx_lt.data(), LPFNT_LOWER, true, false}, {p->Vx_ut.data(), LPFNT_UPPER, true, false}};
            return run_any(p, specs, all_dims(p->m), in, out, batch, true, false, on_device, stream);
        }
        std::vector<SweepSpec> specs{{p->inv_Vx_lt.data(), LPFNT_LOWER, false, false}, {p->inv_Vx_ut.data(), LPFNT_UPPER, false, false}};
        return run_any(p, specs, all_dims(p->m), in, out, batch, true, false, on_device, stream);
    }
    if (!p->inv_Vx_lt.empty()) {
        SweepSpec spec{p->inv_Vx_lt.data(), LPFNT_LOWER, false, false};
        return run_any(p, spec, all_dims(p->m), in, out, batch, true, false, on_device, stream);
    }
    SweepSpec spec{p->Vx_lt.empty() ? nullptr : p->Vx_lt.data(), LPFNT_LOWER, true, false};
    return run_any(p, spec, all_dims(p->m), in, out, batch, true, false, on_device, stream);
}

int lpfnt_ifnt(lpfnt_plan *p, const double *in, double *out, int64_t batch, int arith, int on_device, void *stream) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    const bool fma = arith == LPFNT_ARITH_FMA;
    if (!p->Vx_ut.empty()) {
        // transform.py:596-624: upper mat-vec with U, then lower mat-vec with L
        std::vector<SweepSpec> specs{{p->Vx_ut.data(), LPFNT_UPPER, false, fma}, {p->Vx_lt.empty() ? nullptr : p->Vx_lt.data(), LPFNT_LOWER, false, fma}};
        return run_any(p, specs, all_dims(p->m), in, out, batch, false, true, on_device, stream);
    }
    SweepSpec spec{p->Vx_lt.empty() ? nullptr : p->Vx_lt.data(), LPFNT_LOWER, false, fma};
    return run_any(p, spec, all_dims(p->m), in, out, batch, false, true, on_device, stream);
}

static int check_ik(lpfnt_plan *p, int i, int k) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (i < 0 || i >= p->m) { set_error("Choose a coordinate between i=0 and i=" + std::to_string(p->m - 1) + "."); return LPFNT_EINVAL; }
    if (k < 1 || k > 3) { set_error("Invalid value for k. Please choose 1, 2 or 3."); return LPFNT_EINVAL; }
    return LPFNT_OK;
}

int lpfnt_dx(lpfnt_plan *p, const double *in, double *out, int i, int k, int64_t batch, int on_device, void *stream) {
    int rc = check_ik(p, i, k);
    if (rc) return rc;
    SweepSpec spec{p->Dx_ut[k - 1].empty() ? nullptr : p->Dx_ut[k - 1].data(), LPFNT_UPPER, false, false};
    return run_any(p, spec, std::vector<int>{i}, in, out, batch, false, false, on_device, stream);
}

int lpfnt_dxT(lpfnt_plan *p, const double *in, double *out, int i, int k, int64_t batch, int on_device, void *stream) {
    int rc = check_ik(p, i, k);
    if (rc) return rc;
    SweepSpec spec{p->DxT_lt[k - 1].empty() ? nullptr : p->DxT_lt[k - 1].data(), LPFNT_LOWER, false, false};
    return run_any(p, spec, std::vector<int>{i}, in, out, batch, false, false, on_device, stream);
}

int lpfnt_newton2point(lpfnt_plan *p, const double *coeffs, const double *points, int64_t P, double *values,
                       int on_device, void *stream) {
    if (!p || !coeffs || !values || P < 0 || (P > 0 && !points)) { set_error("eval: bad arguments"); return LPFNT_EINVAL; }
    if (p->nodes.empty()) { set_error("eval: the plan was created without nodes"); return LPFNT_EINVAL; }
    if (P == 0) return LPFNT_OK;
    CK(cudaSetDevice(p->device));
    cudaStream_t s = on_device ? static_cast<cudaStream_t>(stream) : p->stream;
    const double *dc = coeffs, *dp = points;
    double *dv = values;
    // the Chebyshev basis only has the block-staged fast kernel (and the generic fallback)
    const bool fast = !p->force_generic && (p->n + 1 <= kMaxRegT) && p->m <= kEvalMaxM && p->n <= 255 && (p->m == 1 || p->d_line_alpha) &&
                      (!p->chebyshev_eval || (p->use_eval_block && p->d_eval_block && p->d_cpad));
    const size_t cb = size_t(p->N) * sizeof(double), pb = size_t(P) * p->m * sizeof(double), vb = size_t(P) * sizeof(double);
    if (!on_device) {
        int rc = ensure_ws(p, 0, cb + pb);
        if (rc) return rc;
        rc = ensure_ws(p, 1, vb);
        if (rc) return rc;
        double *w = p->ws[0];
        CK(cudaMemcpyAsync(w, coeffs, cb, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w + p->N, points, pb, cudaMemcpyHostToDevice, s));
        dc = w; dp = w + p->N; dv = p->ws[1];
    }
    if (fast) {
        const bool blocked = p->use_eval_block && p->d_eval_block && p->d_cpad && (p->m == 1 || p->d_line_alpha);
        if (blocked) {
            EvalPadArgs pa{dc, p->d_cpad, p->d_line_off, p->d_eval_line, int32_t(p->N1)};
            trace_begin(p, 1100, s);
            cudaError_t e = launch_eval_pad(pa, s);
            trace_end(p, s);
            p->launches += 1;
            if (e != cudaSuccess) { set_error(std::string("eval pad launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
        }
        // grid.x is 32-bit: chunk very large point sets
        const int64_t step = int64_t(1) << 30;
        for (int64_t q0 = 0; q0 < P; q0 += step) {
            const int64_t np = std::min(step, P - q0);
            cudaError_t e;
            if (blocked) {
                EvalBlockArgs a{p->d_cpad, dp + q0 * p->m, dv + q0, p->d_eval_block, p->d_eval_line, p->d_line_alpha, np, p->num_eval_blocks, p->m};
                trace_begin(p, 1200, s);
                e = launch_eval_block(a, p->nodes.data(), p->n + 1, p->chebyshev_eval != 0, s);
            } else {
                EvalArgs a{dc, dp + q0 * p->m, dv + q0, p->d_line_off, p->d_line_alpha, np, int32_t(p->N1), p->m};
                trace_begin(p, 300, s);
                e = launch_eval_horner(a, p->nodes.data(), p->n + 1, s);
            }
            trace_end(p, s);
            p->launches += 1;
            if (e != cudaSuccess) { set_error(std::string("eval launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
        }
    } else {
        if (!p->d_A) {
            if (p->hA.empty()) { set_error("eval: generic path needs the multi-index array"); return LPFNT_EUNSUPPORTED; }
            int rc = upload(&p->d_A, p->hA.data(), p->hA.size(), p);
            if (rc) return rc;
        }
        const int64_t per_point = int64_t(p->m) * (p->n + 1);
        const int64_t step = std::max<int64_t>(128, (int64_t(256) << 20) / (per_point * 8));
        int rc = ensure_ws(p, 2, size_t(std::min(step, P)) * per_point * sizeof(double));
        if (rc) return rc;
        for (int64_t q0 = 0; q0 < P; q0 += step) {
            const int64_t np = std::min(step, P - q0);
            EvalGenericArgs a{dc, dp + q0 * p->m, p->d_nodes, dv + q0, p->d_A, p->ws[2], np, p->N, p->m, p->n, p->chebyshev_eval};
            cudaError_t e = launch_eval_generic(a, s);
            p->launches += 1;
            if (e != cudaSuccess) { set_error(std::string("eval launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
        }
    }
    if (!on_device) {
        CK(cudaMemcpyAsync(values, dv, vb, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    return LPFNT_OK;
}

int lpfnt_eval(lpfnt_plan *p, const double *coeffs, const double *points, int64_t P, double *values, int on_device, void *stream) {
    return lpfnt_newton2point(p, coeffs, points, P, values, on_device, stream);
}

int lpfnt_measure_fp64(int device, int iters, double *fma_per_s) {
    if (!fma_per_s || iters < 1) { set_error("bad arguments"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double *sink = nullptr;
    CK(cudaMalloc(reinterpret_cast<void **>(&sink), 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, nullptr));
        CK(launch_fp64_probe(sink, iters, blocks, nullptr));
        CK(cudaEventRecord(e1, nullptr));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fmas = double(blocks) * 256.0 * 8.0 * double(iters);
        if (rep > 0) best = std::max(best, fmas / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *fma_per_s = best;
    return LPFNT_OK;
}

}  // extern "C"
