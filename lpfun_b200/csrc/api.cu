// C ABI of liblpfnt.so: plan management, operator dispatch, host<->device staging.
// See include/lpfnt.h for the contract of every entry point and the reference call sites
// (src/lpfun/transform.py, src/lpfun/core/molecules.py) each one replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <bitset>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
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
    int2 *groups = nullptr;
    uint8_t *group_t = nullptr, *group_cta_cap = nullptr;
    int32_t num_groups = 0;
    int32_t num_threads = 0, max_lines = 0;
};

int tri_size(int n1) { return n1 * (n1 + 1) / 2; }

}  // namespace

// device copy of a 1-D operator, cached per plan by content: the kernel-layout triangle for every capacity it was used
// with (kernel parameter), and the reference-layout matrix in device memory for the generic kernel
struct OperatorCache {
    std::vector<double> packed;                 // reference packing, as passed by the caller
    int kind = 0;
    std::vector<double> tri[4];                 // kernel layout for capacity 8 / 16 / 24 / 32 (empty until needed)
    double *d_packed = nullptr;                 // reference packing on the device (generic kernel)
};

struct lpfnt_plan {
    int device = 0;
    int m = 0, n = 0;
    int64_t N = 0, N1 = 0;
    int max_line = 0;
    bool has_perm = false;
    // device tables
    int32_t *d_line_off = nullptr, *d_perm = nullptr;
    int32_t *d_line_order = nullptr;
    uint8_t *d_line_alpha = nullptr;
    int2 *d_eval_line = nullptr;
    int2 *d_eval_tline = nullptr;   // k_eval_tree line descriptors (Newton eval, n < 32)
    int4 *d_eval_block = nullptr;
    double *d_cpad = nullptr;       // padded coefficient layout of the eval kernel
    int32_t num_eval_blocks = 0;
    int use_eval_block = 1;
    int use_eval_split = -1;        // few points: split the index trie over CTAs (-1 automatic, 0 off, 1 always)
    std::vector<uint8_t> h_line_alpha;   // host copy (the split levels are derived on first use)
    struct EvalSplit { int L = 0; int32_t num_sub = 0; int32_t *d_l0 = nullptr; uint8_t *d_alpha = nullptr; };
    std::vector<EvalSplit> eval_splits;  // one per level L already built
    std::vector<int64_t> eval_split_counts;  // subtrees per level L = 1 .. m-1 (filled on first use)
    int64_t eval_padded = 0;        // doubles of the padded coefficient layout (d_cpad is allocated by the first eval that needs it)
    int64_t *d_A = nullptr;       // only for the generic eval fallback (uploaded lazily)
    double *d_nodes = nullptr;
    DeviceDim dd[kMaxDim + 1];
    std::vector<int64_t> hA;      // kept only when the generic eval may need it
    // host copies of the 1-D operators (reference packing)
    std::vector<double> nodes, Vx_lt, inv_Vx_lt, Dx_ut[3], DxT_lt[3];
    // LU-factorised (non-triangular) V, e.g. the Chebyshev basis: upper factors; chebyshev_eval selects the
    // T_k basis in eval (transform.py:292-302, utils.py:165-195)
    std::vector<double> Vx_ut, inv_Vx_ut;
    int chebyshev_eval = 0;
    std::vector<OperatorCache *> ops;
    // workspaces (grown on demand): two lanes x {staged input, staged output, permutation scratch}
    double *ws[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t ws_bytes[6] = {0, 0, 0, 0, 0, 0};
    int dbg_uploads = 0;            // LPFNT_DEBUG_BYTES: running number of the table uploads
    cudaStream_t stream = nullptr, stream2 = nullptr;  // used by the host-pointer paths (two lanes)
    // pageable host arrays (the reference's plain NumPy contract) are staged through pinned buffers by several host threads
    double *hstage[4] = {nullptr, nullptr, nullptr, nullptr};  // lane 0 in / out, lane 1 in / out
    size_t hstage_bytes[4] = {0, 0, 0, 0};
    cudaEvent_t lane_done[2] = {nullptr, nullptr};
    int host_threads = 0;                    // staging copy threads (0 = automatic)
    int64_t launches = 0;
    int64_t table_bytes = 0;
    int force_generic = 0;
    int dx_descending = 0;                   // dx folds its rows from the last term down (the reference on p = inf sets with m > 1)
    // whole-function kernel (m = 2, 3): tables of HostIndex::whole on the device, one set for the exact forms
    // (index 0) and one for the fused-multiply-add forms (index 1): their best CTA shapes differ
    struct Whole {
        WholeHost h;                         // host copy (items / seo dropped after the upload)
        uint32_t *d_items = nullptr;
        uint16_t *d_seo = nullptr;
        int32_t total_items = 0, seo_len = 0;
        size_t smem_bytes = 0;
    } whole[2];
    struct WholeCfg { int threads = -1, fibres = -1, order = kWholeDefaultOrder, rows = -1; } wcfg[2];
    uint16_t *d_perm16 = nullptr;
    // push region: every rank's mapping of a symmetric output buffer (lpfnt_plan_set_push)
    struct Push { const char *local = nullptr; int64_t bytes = 0; int n = 0; char *peer[8] = {nullptr}; char *mc = nullptr; } push;
    // split whole-function kernel (two units per function, two CTAs per SM; m = 3, lower mat-vec)
    struct WholeSplit {
        WholeSplitHost h;                    // host copy (items / seo dropped after the upload)
        uint32_t *d_items = nullptr;
        uint16_t *d_seo = nullptr;
        double *d_scratch = nullptr;
        int32_t total_items = 0, seo_len = 0, buf_elems = 0;
        size_t smem_bytes = 0;
    } wsplit;
    int use_whole_split = 0;                 // 0 off (default: measured slower), 1 whenever the tables exist, -1 automatic
    int whole_split_c = 0, whole_split_order = kWholeDefaultOrder, whole_split_skew = 0;
    std::bitset<1024> whole_attr_set;         // per kernel instantiation: dynamic shared memory opted in on this device
    int use_whole = -1;                      // -1 auto (batches of >= 8 functions), 0 off, 1 always
    HostIndex *hx = nullptr;                 // kept for plans that can run the whole-function kernel (re-build of its tables)
    // slab kernel (coordinates 0 and 1 in one pass, any m >= 2; see k_slab)
    struct Slabs {
        bool ok = false;
        int32_t n_slabs = 0, max_t = 0;
        SlabDesc *d_desc = nullptr;
        int2 *d_item0 = nullptr, *d_item1 = nullptr;
    } slabs;
    int use_slab = 1;
    int slab_ctas = 4;                       // CTAs per SM the persistent slab grid is sized for (capacities <= 24)
    int use_cta_cap = 1;                     // k_sweep_groups: per-CTA row capacity (8 / 16 / MAXT)
    int64_t host_chunk_bytes = 0;            // host-pointer calls: bytes per pipelined chunk (H2D | sweeps | D2H on two lanes); 0 = automatic
    long long *d_timeline = nullptr;         // development: clock64 stamps of k_whole (option "timeline")
    int num_sms = 148;
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
    static const bool dbg = std::getenv("LPFNT_DEBUG_BYTES") != nullptr;
    if (dbg && count * sizeof(T) >= (size_t(1) << 20)) std::fprintf(stderr, "lpfnt upload #%d: %.1f MB (tables %.1f MB)\n", ++p->dbg_uploads, count * sizeof(T) / 1e6, p->table_bytes / 1e6);
    else if (dbg) ++p->dbg_uploads;
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

// The operator `packed` (host, reference packing) as the plan knows it: looked up by CONTENT, so a caller may pass the
// same matrix from any buffer; re-packing and the upload for the generic kernel happen once per plan, not per call.
OperatorCache *find_operator(lpfnt_plan *p, const double *packed, int kind) {
    const size_t cnt = size_t(tri_size(p->n + 1));
    for (OperatorCache *oc : p->ops)
        if (oc->kind == kind && std::memcmp(oc->packed.data(), packed, cnt * sizeof(double)) == 0) return oc;
    if (p->ops.size() >= 64) {  // a caller streaming ever-changing matrices: recycle the oldest entry
        OperatorCache *old = p->ops.front();
        p->ops.erase(p->ops.begin());
        if (old->d_packed) { cudaDeviceSynchronize(); cudaFree(old->d_packed); }
        delete old;
    }
    OperatorCache *oc = new OperatorCache();
    oc->packed.assign(packed, packed + cnt);
    oc->kind = kind;
    p->ops.push_back(oc);
    return oc;
}
const double *operator_tri(lpfnt_plan *p, OperatorCache *oc, int cap) {
    std::vector<double> &t = oc->tri[cap / 8 - 1];
    if (t.empty()) repack(oc->packed.data(), p->n + 1, oc->kind, cap, t);
    return t.data();
}
int operator_device(lpfnt_plan *p, OperatorCache *oc, const double **d) {
    if (!oc->d_packed) {
        CK(cudaMalloc(reinterpret_cast<void **>(&oc->d_packed), oc->packed.size() * sizeof(double)));
        CK(cudaMemcpy(oc->d_packed, oc->packed.data(), oc->packed.size() * sizeof(double), cudaMemcpyHostToDevice));
        p->table_bytes += int64_t(oc->packed.size() * sizeof(double));
    }
    *d = oc->d_packed;
    return LPFNT_OK;
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
    OperatorCache *op = nullptr;  // resolved once per call (run_any)
    bool desc = false;            // upper mat-vec folded from the last term down (dx on p = inf sets, molecules.py:290-294)
};

int sweep_mode(const SweepSpec &s) {
    if (s.kind == LPFNT_LOWER) return s.solve ? kLowerSolve : kLowerMatvec;
    return s.solve ? kUpperSolve : (s.desc ? kUpperMatvecDesc : kUpperMatvec);
}

#define LAUNCH_CAP(cap, call8, call16, call24, call32) \
    switch (cap) { case 8: e = call8; break; case 16: e = call16; break; case 24: e = call24; break; default: e = call32; break; }

int launch_slab_pass(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s);

// One sweep along coordinate h: src -> dst (device pointers, may alias when no perm is used).
int launch_dim(lpfnt_plan *p, const SweepSpec &spec, int h, bool fuse0, const double *src, double *dst, int64_t batch,
               bool gather, bool scatter, cudaStream_t s) {
    const int n1 = p->n + 1;
    int mode = sweep_mode(spec);
    if (mode == kUpperSolve && h >= 1) mode = kUpperSolveDesc;  // the reference's fold order there (atoms.py:391-402)
    if (fuse0) {
        if (h != 1) { set_error("internal: only coordinate 1 fuses with coordinate 0"); return LPFNT_EINVAL; }
        return launch_slab_pass(p, spec, src, dst, batch, gather, scatter, s);
    }
    const int len = (h == 0) ? p->max_line : p->dd[h].max_lines;
    const int cap = p->force_generic ? 0 : pick_cap(std::max(len, 1));
    if (cap && h >= 1 && !p->dd[h].groups) { set_error("internal: lane-group tables missing"); return LPFNT_EINVAL; }
    const double *tri = cap ? operator_tri(p, spec.op, cap) : nullptr;
    const double *d_mat = nullptr;
    if (!cap) { int rc = operator_device(p, spec.op, &d_mat); if (rc) return rc; }
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = int(std::min<int64_t>(65535, batch - b0));
        const double *in = src + b0 * p->N;
        double *out = dst + b0 * p->N;
        cudaError_t e = cudaSuccess;
        trace_begin(p, (cap == 0 ? 200 : (h == 0 ? 0 : 100)) + h, s);
        if (cap == 0) {
            GenericArgs a{};
            a.in = in; a.out = out; a.mat = d_mat; a.line_off = p->d_line_off;
            a.perm_in = gather ? p->d_perm : nullptr;
            a.perm_out = scatter ? p->d_perm : nullptr;
            a.stride = p->N; a.n1 = n1; a.mode = mode; a.fma = spec.fma ? 1 : 0;
            if (h == 0) { a.ent = nullptr; a.num_threads = int32_t(p->N1); }
            else {
                const DeviceDim &d = p->dd[h];
                if (!d.thr_fib) { set_error("force_generic: the fibre CSR of this plan was not uploaded (set larger than 2^20 elements)"); return LPFNT_EUNSUPPORTED; }
                a.thr_fib = d.thr_fib; a.thr_start = d.thr_start; a.fib_ptr = d.fib_ptr; a.ent = d.ent;
                a.num_threads = d.num_threads;
            }
            e = launch_generic_sweep(a, nb, s);
        } else if (h == 0) {
            LineSweepArgs a{in, out, p->d_line_off, gather ? p->d_perm : nullptr, scatter ? p->d_perm : nullptr, p->d_line_order, p->N, int32_t(p->N1)};
            LAUNCH_CAP(cap, launch_line_sweep<8>(mode, spec.fma, a, tri, nb, s), launch_line_sweep<16>(mode, spec.fma, a, tri, nb, s),
                       launch_line_sweep<24>(mode, spec.fma, a, tri, nb, s), launch_line_sweep<32>(mode, spec.fma, a, tri, nb, s))
        } else {
            if (gather) { set_error("internal: gather is only fused into the coordinate-0 sweep"); return LPFNT_EINVAL; }
            const DeviceDim &d = p->dd[h];
            GroupSweepArgs ga{in, out, d.groups, d.group_t, d.ent, scatter ? p->d_perm : nullptr, p->N, d.num_groups, p->use_cta_cap ? d.group_cta_cap : nullptr};
            LAUNCH_CAP(cap, launch_group_sweep<8>(mode, spec.fma, ga, tri, nb, s), launch_group_sweep<16>(mode, spec.fma, ga, tri, nb, s),
                       launch_group_sweep<24>(mode, spec.fma, ga, tri, nb, s), launch_group_sweep<32>(mode, spec.fma, ga, tri, nb, s))
        }
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

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
        // result stores, fewer bank conflicts) at ~10 % less uniform lengths
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
    const double *tri = operator_tri(p, spec.op, cap);
    SlabArgs a{};
    a.slab = p->slabs.d_desc; a.item0 = p->slabs.d_item0; a.item1 = p->slabs.d_item1; a.line_off = p->d_line_off;
    a.perm_in = gather ? p->d_perm : nullptr;
    a.perm_out = scatter ? p->d_perm : nullptr;
    a.stride = p->N; a.n_slabs = p->slabs.n_slabs;
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 24)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 24, batch - b0));
        a.in = src + b0 * p->N; a.out = dst + b0 * p->N; a.batch = nb;
        const int64_t units = int64_t(a.n_slabs) * nb;
        const int grid = int(std::min<int64_t>(units, int64_t(p->num_sms) * (cap <= 24 ? p->slab_ctas : 3)));
        cudaError_t e = cudaSuccess;
        trace_begin(p, 1601, s);
        LAUNCH_CAP(cap, launch_slab<8>(mode, spec.fma, a, tri, grid, s), launch_slab<16>(mode, spec.fma, a, tri, grid, s),
                   launch_slab<24>(mode, spec.fma, a, tri, grid, s), launch_slab<32>(mode, spec.fma, a, tri, grid, s))
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("slab kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    }
    return LPFNT_OK;
}

// ---- whole-function kernel (k_whole): tables, eligibility, launch ----
void free_whole(lpfnt_plan *p, int which) {
    lpfnt_plan::Whole &W = p->whole[which];
    cudaFree(W.d_items); cudaFree(W.d_seo);
    W.d_items = nullptr; W.d_seo = nullptr;
    W.h = WholeHost();
}

// CTA shape: two CTAs of 384 threads (one fibre per thread) per SM when two functions (plus tables) fit into shared memory;
// else one CTA per SM: the fused forms with 384 threads and TWO fibres each, folded side by side (every matrix entry feeds
// two DFMA: the constant path, one LDCU.64 per ~4 cycles and SM sub-partition, is half as loaded: ifnt 0.53 -> 0.49 ms at
// d = 3, n = 30, B = 4096), the exact forms with 768 threads and one fibre each.  (A third form -- 384 threads with a long
// and a short item each, one after the other: equal work per warp, 88 % of the FP64 pipe in the triangle phases, but load /
// store phases of 12 instead of 24 warps that grow by more, 0.69 vs 0.61 ms -- was measured and removed)
int upload_whole(lpfnt_plan *p, int which) {
    lpfnt_plan::Whole &W = p->whole[which];
    const lpfnt_plan::WholeCfg &cfg = p->wcfg[which];
    free_whole(p, which);
    if (!p->hx) return LPFNT_OK;
    for (int pass = 0; pass < 2; ++pass) {
        int tg = cfg.threads, nf = cfg.fibres;
        if (pass == 0) { if (tg <= 0) tg = 384; if (nf <= 0) nf = 1; }
        else { if (nf <= 0) nf = (tg == 768) ? 1 : (which == 1 ? 2 : 1); if (tg <= 0) tg = nf == 1 ? 768 : 384; }
        if (nf >= 2) tg = 384;  // two fibres per thread: 168 registers, i.e. 384 threads
        p->hx->build_whole(tg, cfg.order, nf);
        const WholeHost &h = p->hx->whole;
        if (!h.ok) return LPFNT_OK;
        W.smem_bytes = whole_smem_bytes(int(p->N), int(h.items.size()), int(h.seo.size()), p->has_perm);
        const bool explicit_shape = cfg.threads > 0 && cfg.fibres > 0;
        if (explicit_shape || pass == 1 || (tg == 384 && nf == 1 && W.smem_bytes <= size_t(kWholeSmemPerCta))) break;
    }
    W.h = p->hx->whole;
    const bool two_ctas = W.h.threads == 384 && W.h.fibres == 1 && W.smem_bytes <= size_t(kWholeSmemPerCta);
    if (W.smem_bytes > size_t(two_ctas ? kWholeSmemPerCta : kWholeSmemLimit)) { W.h = WholeHost(); return LPFNT_OK; }
    int rc;
    if ((rc = upload(&W.d_items, W.h.items.data(), W.h.items.size(), p))) return rc;
    if ((rc = upload(&W.d_seo, W.h.seo.data(), W.h.seo.size(), p))) return rc;
    W.total_items = int32_t(W.h.items.size());
    W.seo_len = int32_t(W.h.seo.size());
    W.h.items.clear(); W.h.items.shrink_to_fit();
    W.h.seo.clear(); W.h.seo.shrink_to_fit();
    return LPFNT_OK;
}

bool whole_ok(const lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, int64_t batch) {
    const lpfnt_plan::Whole &W = p->whole[spec.fma ? 1 : 0];
    if (!W.h.ok || p->force_generic || int(dims.size()) != p->m) return false;
    // automatic: enough functions to fill a few SMs (a single function would run on ONE SM) and enough items to keep
    // most of a CTA busy
    const bool want = p->use_whole == 1 || (p->use_whole == -1 && batch >= 8 && 3 * W.h.max_items >= 2 * W.h.threads);
    if (!want) return false;
    for (int q = 0; q < p->m; ++q) if (dims[q] != q) return false;
    if ((spec.solve || spec.desc) && spec.kind == LPFNT_UPPER) return false;  // fold order changes with the coordinate / descending fold
    if (W.h.fibres == 2 && sweep_mode(spec) != kLowerMatvec) return false;  // two fibres per thread: lower mat-vec only
    if (reinterpret_cast<uintptr_t>(in) & 15) return false;    // bulk copies need a 16-byte aligned batch
    return true;
}

int launch_whole_fn(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s) {
    const int which = spec.fma ? 1 : 0;
    lpfnt_plan::Whole &W = p->whole[which];
    const int mode = sweep_mode(spec);
    const int cap = pick_cap(std::max(W.h.max_t, 1));
    const double *tri = operator_tri(p, spec.op, cap);
    WholeArgs a{};
    a.items = W.d_items; a.seo = W.d_seo; a.perm16 = p->d_perm16;
    a.total_items = W.total_items; a.seo_len = W.seo_len;
    for (int h = 0; h < 3; ++h) { a.item_off[h] = W.h.item_off[h]; a.n_items[h] = W.h.n_items[h]; }
    a.gather = gather ? 1 : 0; a.scatter = scatter ? 1 : 0;
    a.n_elem = int32_t(p->N); a.m = p->m;
    a.dbg = p->d_timeline;
    // results that land inside the registered symmetric region are stored into every rank's copy of it
    const char *dstc = reinterpret_cast<const char *>(dst);
    const bool push = p->push.n > 0 && dstc >= p->push.local && dstc + size_t(batch) * p->N * sizeof(double) <= p->push.local + p->push.bytes;
    const bool timeline = p->d_timeline != nullptr && !push;
    const int rows = 2;  // rows folded side by side
    const int modekind = mode == kLowerMatvec ? (spec.fma ? 1 : 0) : mode == kUpperMatvec ? (spec.fma ? 3 : 2) : 4;
    const int shape = (W.h.threads == 384 ? 0 : 1) + (W.h.fibres == 2 ? 2 : 0);  // < 8
    const int key = ((cap / 8 - 1) * 5 + modekind) * 8 + shape + (timeline ? 160 : 0) + (push ? 512 : 0);  // < 320 (+ 512 for the push variants)
    const int fibres_code = W.h.fibres == 2 ? 2 : 1;
    const bool set_attr = !p->whole_attr_set[key];
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 30)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 30, batch - b0));
        a.in = src + b0 * p->N; a.out = dst + b0 * p->N; a.batch = nb;
        if (push) {
            const int64_t off = (dstc - p->push.local) + b0 * p->N * int64_t(sizeof(double));
            a.n_peers = p->push.n;
            for (int r = 0; r < p->push.n; ++r) a.peer_out[r] = reinterpret_cast<double *>(p->push.peer[r] + off);
            a.mc_out = p->push.mc ? reinterpret_cast<double *>(p->push.mc + off) : nullptr;
        }
        const int per_sm = (W.h.threads == 384 && W.h.fibres == 1 && W.smem_bytes <= size_t(kWholeSmemPerCta)) ? 2 : 1;
        const int grid = int(std::min<int64_t>(nb, int64_t(per_sm) * p->num_sms));
        cudaError_t e = cudaSuccess;
        trace_begin(p, 1400, s);
        LAUNCH_CAP(cap, launch_whole<8>(mode, spec.fma, W.h.threads, rows, fibres_code, timeline, a, tri, grid, W.smem_bytes, set_attr, s),
                   launch_whole<16>(mode, spec.fma, W.h.threads, rows, fibres_code, timeline, a, tri, grid, W.smem_bytes, set_attr, s),
                   launch_whole<24>(mode, spec.fma, W.h.threads, rows, fibres_code, timeline, a, tri, grid, W.smem_bytes, set_attr, s),
                   launch_whole<32>(mode, spec.fma, W.h.threads, rows, fibres_code, timeline, a, tri, grid, W.smem_bytes, set_attr, s))
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("whole-function kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
        p->whole_attr_set[key] = true;
    }
    return LPFNT_OK;
}


// ---- split whole-function kernel (k_whole_split) ----
void free_wsplit(lpfnt_plan *p) {
    lpfnt_plan::WholeSplit &W = p->wsplit;
    cudaFree(W.d_items); cudaFree(W.d_seo); cudaFree(W.d_scratch);
    W.d_items = nullptr; W.d_seo = nullptr; W.d_scratch = nullptr;
    W.h = WholeSplitHost();
}

int upload_wsplit(lpfnt_plan *p) {
    lpfnt_plan::WholeSplit &W = p->wsplit;
    free_wsplit(p);
    for (int k = 400; k < 408; ++k) p->whole_attr_set[k] = false;  // the shared-memory size changes with the tables
    if (!p->hx || p->m != 3) return LPFNT_OK;
    // the split plane: as given, else the one for which the larger part is smallest among those that fit twice into an SM
    int best_c = 0;
    if (p->whole_split_c > 0) best_c = p->whole_split_c;
    else {
        int64_t best = -1;
        for (int c = 1; c <= 16; ++c) {
            p->hx->build_whole_split(c, p->whole_split_order);
            const WholeSplitHost &h = p->hx->wsplit;
            if (!h.ok) continue;
            const int be = (std::max(h.ne[0], h.ne[1]) + 5) & ~1;
            const size_t sm = whole_split_smem_bytes(be, int(h.items.size()), int(h.seo.size()), int(p->N), p->has_perm);
            if (sm > size_t(kWholeSmemPerCta)) continue;
            // prefer one round per sweep for coordinates 0 and 1, then the smaller buffer
            const int64_t score = int64_t(h.n_items[0][0] + h.n_items[0][1] + h.n_items[1][0] + h.n_items[1][1]) * 100000 + be;
            if (best < 0 || score < best) { best = score; best_c = c; }
        }
    }
    if (best_c <= 0) return LPFNT_OK;
    p->hx->build_whole_split(best_c, p->whole_split_order);
    if (!p->hx->wsplit.ok) return LPFNT_OK;
    W.h = p->hx->wsplit;
    W.buf_elems = (std::max(W.h.ne[0], W.h.ne[1]) + 5) & ~1;
    W.smem_bytes = whole_split_smem_bytes(W.buf_elems, int(W.h.items.size()), int(W.h.seo.size()), int(p->N), p->has_perm);
    if (W.smem_bytes > size_t(kWholeSmemPerCta)) { W.h = WholeSplitHost(); return LPFNT_OK; }
    int rc;
    if ((rc = upload(&W.d_items, W.h.items.data(), W.h.items.size(), p))) return rc;
    if ((rc = upload(&W.d_seo, W.h.seo.data(), W.h.seo.size(), p))) return rc;
    const size_t sc = size_t(2) * p->num_sms * size_t(W.h.C) * size_t(std::max(W.h.ncolq, 1));
    CK(cudaMalloc(reinterpret_cast<void **>(&W.d_scratch), sc * sizeof(double)));
    p->table_bytes += int64_t(sc * sizeof(double));
    W.total_items = int32_t(W.h.items.size());
    W.seo_len = int32_t(W.h.seo.size());
    W.h.items.clear(); W.h.items.shrink_to_fit();
    W.h.seo.clear(); W.h.seo.shrink_to_fit();
    return LPFNT_OK;
}

bool wsplit_ok(const lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, int64_t batch) {
    const lpfnt_plan::WholeSplit &W = p->wsplit;
    if (!W.h.ok || p->force_generic || p->use_whole == 0 || p->use_whole_split == 0 || int(dims.size()) != 3) return false;
    for (int q = 0; q < 3; ++q) if (dims[q] != q) return false;
    if (sweep_mode(spec) != kLowerMatvec) return false;
    if (reinterpret_cast<uintptr_t>(in) & 15) return false;
    if (pick_cap(std::max(W.h.max_t, 1)) < 32) return false;  // smaller functions (n <= 23: N <= 7.3 K) fit twice into an SM unsplit
    // automatic: only when the unsplit kernel is limited to one CTA per SM, and for batches that fill the GPU
    if (p->use_whole_split == 1) return true;
    const lpfnt_plan::Whole &U = p->whole[spec.fma ? 1 : 0];
    const bool unsplit_two_ctas = U.h.ok && U.h.threads == 384 && U.h.fibres == 1 && U.smem_bytes <= size_t(kWholeSmemPerCta);
    return p->use_whole_split == -1 && p->use_whole != 0 && !unsplit_two_ctas && batch >= 2 * p->num_sms;
}

int launch_wsplit_fn(lpfnt_plan *p, const SweepSpec &spec, const double *src, double *dst, int64_t batch, bool gather, bool scatter, cudaStream_t s) {
    lpfnt_plan::WholeSplit &W = p->wsplit;
    const int cap = pick_cap(std::max(W.h.max_t, 1));
    const double *tri = operator_tri(p, spec.op, cap);
    WholeSplitArgs a{};
    a.items = W.d_items; a.seo = W.d_seo; a.perm16 = p->d_perm16; a.scratch = W.d_scratch;
    a.dbg = p->d_timeline;
    for (int q = 0; q < 2; ++q) {
        a.e0[q] = W.h.e0[q]; a.ne[q] = W.h.ne[q];
        for (int h = 0; h < 3; ++h) { a.item_off[q][h] = W.h.item_off[q][h]; a.n_items[q][h] = W.h.n_items[q][h]; }
    }
    for (int q = 0; q < 32; ++q) { a.cb[q] = W.h.cb[q]; a.rq[q] = W.h.rq[q]; }
    a.total_items = W.total_items; a.seo_len = W.seo_len; a.ncolq = std::max(W.h.ncolq, 1); a.C = W.h.C;
    a.buf_elems = W.buf_elems;
    a.gather = gather ? 1 : 0; a.scatter = scatter ? 1 : 0;
    a.n_elem = int32_t(p->N);
    a.skew = p->whole_split_skew;
    const bool timeline = p->d_timeline != nullptr;
    const int key = 400 + (spec.fma ? 1 : 0) + (timeline ? 2 : 0);
    const bool set_attr = !p->whole_attr_set[key];
    for (int64_t b0 = 0; b0 < batch; b0 += (int64_t(1) << 30)) {
        const int nb = int(std::min<int64_t>(int64_t(1) << 30, batch - b0));
        a.in = src + b0 * p->N; a.out = dst + b0 * p->N; a.batch = nb;
        const int grid = int(std::min<int64_t>(nb, int64_t(2) * p->num_sms));
        cudaError_t e = cudaSuccess;
        trace_begin(p, 1500, s);
        e = cap == 32 ? launch_whole_split<32>(spec.fma, timeline, a, tri, grid, W.smem_bytes, set_attr, s) : cudaErrorInvalidValue;
        trace_end(p, s);
        p->launches += 1;
        if (e != cudaSuccess) { set_error(std::string("split whole-function kernel launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
        p->whole_attr_set[key] = true;
    }
    return LPFNT_OK;
}

// Device-resident pipeline: sweeps along `dims` (ascending), optional colex gather on the way
// in and scatter on the way out.  in/out may alias.
//
// Pass plan: the whole-function kernel when it applies (one pass); else the slab kernel takes coordinates 0 and 1 in one
// pass and every remaining coordinate is one pass.  Passes 2.. run in place.
struct Pass { int h; bool fuse0; };

int run_device(lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, double *out,
               int64_t batch, bool gather, bool scatter, cudaStream_t s, int lane) {
    const int mid_slot = 3 * lane + 2;
    gather = gather && p->has_perm;
    scatter = scatter && p->has_perm;
    if (dims.empty()) { set_error("no sweep dimensions"); return LPFNT_EINVAL; }
    if (gather && dims[0] != 0) { set_error("gather requires a sweep along coordinate 0"); return LPFNT_EINVAL; }
    {
        const char *oc = reinterpret_cast<const char *>(out);
        const bool in_push = p->push.n > 0 && oc >= p->push.local && oc < p->push.local + p->push.bytes;
        if (in_push) {
            if (!whole_ok(p, spec, dims, in, batch)) { set_error("output lies in the push region, but this call does not run on the whole-function kernel (the only kernel that stores into the peers' buffers)"); return LPFNT_EUNSUPPORTED; }
            return launch_whole_fn(p, spec, in, out, batch, gather, scatter, s);
        }
    }
    if (wsplit_ok(p, spec, dims, in, batch)) return launch_wsplit_fn(p, spec, in, out, batch, gather, scatter, s);
    if (whole_ok(p, spec, dims, in, batch)) {
        // in-place is safe: a function is complete in shared memory / registers before its first store
        return launch_whole_fn(p, spec, in, out, batch, gather, scatter, s);
    }
    std::vector<Pass> passes;
    {
        size_t q = 0;
        const bool can_slab = p->use_slab && p->slabs.ok && !p->force_generic && p->m >= 2 &&
                              pick_cap(std::max(p->slabs.max_t, 1)) > 0 && !((spec.solve || spec.desc) && spec.kind == LPFNT_UPPER);
        if (dims.size() >= 2 && dims[0] == 0 && dims[1] == 1 && can_slab) { passes.push_back({1, true}); q = 2; }
        for (; q < dims.size(); ++q) passes.push_back({dims[q], false});
    }
    const int np = int(passes.size());
    const size_t fbytes = size_t(p->N) * sizeof(double);
    const bool permuting = gather || scatter;
    // a permuting pass must not run in place; with >= 2 passes the middle buffer absorbs that
    const bool need_mid = (np == 1) ? (permuting && in == out) : (scatter || (permuting && in == out));
    double *mid_base = nullptr;
    if (need_mid) { int rc = ensure_ws(p, mid_slot, size_t(batch) * fbytes); if (rc) return rc; mid_base = p->ws[mid_slot]; }
    if (np == 1) {
        double *dst = need_mid ? mid_base : out;
        int rc = launch_dim(p, spec, passes[0].h, passes[0].fuse0, in, dst, batch, gather, scatter, s);
        if (rc) return rc;
        if (dst != out) CK(cudaMemcpyAsync(out, dst, size_t(batch) * fbytes, cudaMemcpyDeviceToDevice, s));
        return LPFNT_OK;
    }
    double *mid = need_mid ? mid_base : out;
    for (int q = 0; q < np; ++q) {
        const bool first = (q == 0), last = (q == np - 1);
        int rc = launch_dim(p, spec, passes[q].h, passes[q].fuse0, first ? in : mid, last ? out : mid, batch,
                            first && gather, last && scatter, s);
        if (rc) return rc;
    }
    return LPFNT_OK;
}

// Sequence of sweeps (one spec = one triangular operator applied along `dims`); gather rides on the first,
// scatter on the last.
int run_specs(lpfnt_plan *p, const std::vector<SweepSpec> &specs, const std::vector<int> &dims, const double *in, double *out,
              int64_t batch, bool gather, bool scatter, cudaStream_t s, int lane) {
    const int ns = int(specs.size());
    for (int i = 0; i < ns; ++i) {
        int rc = run_device(p, specs[i], dims, i == 0 ? in : out, out, batch, gather && i == 0, scatter && i == ns - 1, s, lane);
        if (rc) return rc;
    }
    return LPFNT_OK;
}


// true for ordinary (pageable, unregistered) host memory
bool is_pageable(const void *ptr) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// memcpy split over a few host threads: one core moves ~10 GB/s, PCIe Gen5 takes 55 GB/s per direction
void parallel_copy(void *dst, const void *src, size_t bytes, int threads) {
    if (threads <= 1 || bytes < (size_t(1) << 20)) { std::memcpy(dst, src, bytes); return; }
    const size_t piece = ((bytes / size_t(threads)) + 4095) & ~size_t(4095);
    std::vector<std::thread> th;
    for (size_t o = piece; o < bytes; o += piece)
        th.emplace_back([=]() { std::memcpy(static_cast<char *>(dst) + o, static_cast<const char *>(src) + o, std::min(piece, bytes - o)); });
    std::memcpy(dst, src, std::min(piece, bytes));
    for (auto &t : th) t.join();
}

int ensure_hstage(lpfnt_plan *p, int slot, size_t bytes) {
    if (p->hstage_bytes[slot] >= bytes) return LPFNT_OK;
    if (p->hstage[slot]) CK(cudaFreeHost(p->hstage[slot]));
    p->hstage[slot] = nullptr; p->hstage_bytes[slot] = 0;
    CK(cudaHostAlloc(reinterpret_cast<void **>(&p->hstage[slot]), bytes, cudaHostAllocPortable));
    p->hstage_bytes[slot] = bytes;
    return LPFNT_OK;
}

// Host-pointer front end: stage through the lanes' workspaces in chunks of functions.
int run_any(lpfnt_plan *p, std::vector<SweepSpec> specs, const std::vector<int> &dims, const double *in, double *out,
            int64_t batch, bool gather, bool scatter, int on_device, void *stream) {
    if (!p || !in || !out || batch < 1) { set_error("null pointer or empty batch"); return LPFNT_EINVAL; }
    for (const SweepSpec &sp : specs)
        if (!sp.packed) { set_error("the plan was created without the matrix this operation needs"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(p->device));
    for (SweepSpec &sp : specs) sp.op = find_operator(p, sp.packed, sp.kind);
    if (on_device) return run_specs(p, specs, dims, in, out, batch, gather, scatter, static_cast<cudaStream_t>(stream), 0);
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
    // Pageable host memory (plain NumPy arrays): cudaMemcpyAsync would stage it through the driver's own bounce buffer with
    // one thread (measured 6-7 GB/s end to end); instead the chunks go through pinned buffers of the plan, copied by a few
    // host threads while the other lane's chunk is on the GPU.  Pinned / registered memory is passed to the DMA engines as is.
    const size_t total = size_t(batch) * fbytes;
    const bool stage_in = total >= (size_t(4) << 20) && is_pageable(in);
    const bool stage_out = total >= (size_t(4) << 20) && is_pageable(out);
    const int hthreads = p->host_threads > 0 ? p->host_threads : int(std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2)));  // measured: ~linear up to 8 threads (0.43 -> 1.5 GDOF/s at d=3, n=30)
    if (stage_in || stage_out) {
        for (int l = 0; l < lanes; ++l) {
            int rc = LPFNT_OK;
            if (stage_in) rc = ensure_hstage(p, 2 * l, size_t(chunk) * fbytes);
            if (!rc && stage_out) rc = ensure_hstage(p, 2 * l + 1, size_t(chunk) * fbytes);
            if (rc) return rc;
            if (!p->lane_done[l]) CK(cudaEventCreateWithFlags(&p->lane_done[l], cudaEventDisableTiming));
        }
    }
    int64_t pend_b0[2] = {-1, -1}, pend_nb[2] = {0, 0};  // chunk whose result still sits in the lane's pinned out buffer
    auto flush_lane = [&](int l) -> int {
        if (pend_b0[l] < 0) return LPFNT_OK;
        CK(cudaEventSynchronize(p->lane_done[l]));
        parallel_copy(out + pend_b0[l] * p->N, p->hstage[2 * l + 1], size_t(pend_nb[l]) * fbytes, hthreads);
        pend_b0[l] = -1;
        return LPFNT_OK;
    };
    int i = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += chunk, ++i) {
        const int64_t nb = std::min(chunk, batch - b0);
        const int l = (lanes == 2) ? (i & 1) : 0;
        cudaStream_t s = l ? p->stream2 : p->stream;
        if (stage_out) { int rc = flush_lane(l); if (rc) return rc; }
        else if (stage_in && i >= lanes) CK(cudaEventSynchronize(p->lane_done[l]));  // the lane's pinned in buffer is free again
        const double *src = in + b0 * p->N;
        if (stage_in) {
            parallel_copy(p->hstage[2 * l], src, size_t(nb) * fbytes, hthreads);
            src = p->hstage[2 * l];
        }
        CK(cudaMemcpyAsync(p->ws[3 * l], src, size_t(nb) * fbytes, cudaMemcpyHostToDevice, s));
        int rc = run_specs(p, specs, dims, p->ws[3 * l], p->ws[3 * l + 1], nb, gather, scatter, s, l);
        if (rc) return rc;
        double *dst = stage_out ? p->hstage[2 * l + 1] : out + b0 * p->N;
        CK(cudaMemcpyAsync(dst, p->ws[3 * l + 1], size_t(nb) * fbytes, cudaMemcpyDeviceToHost, s));
        if (stage_in || stage_out) CK(cudaEventRecord(p->lane_done[l], s));
        if (stage_out) { pend_b0[l] = b0; pend_nb[l] = nb; }
    }
    if (stage_out) {
        // drain in issue order
        const int first = (lanes == 2) ? (i & 1) : 0;
        for (int q = 0; q < lanes; ++q) { int rc = flush_lane((first + q) % lanes); if (rc) return rc; }
    }
    CK(cudaStreamSynchronize(p->stream));
    if (lanes == 2) CK(cudaStreamSynchronize(p->stream2));
    return LPFNT_OK;
}

int run_any(lpfnt_plan *p, const SweepSpec &spec, const std::vector<int> &dims, const double *in, double *out,
            int64_t batch, bool gather, bool scatter, int on_device, void *stream) {
    return run_any(p, std::vector<SweepSpec>{spec}, dims, in, out, batch, gather, scatter, on_device, stream);
}


// Subtrees of the index trie below its top L levels (see k_eval_sub): runs of lines with equal (a_{m-L}, .., a_{m-1}).
int eval_split_level(lpfnt_plan *p, int L, const lpfnt_plan::EvalSplit **out) {
    for (const auto &es : p->eval_splits)
        if (es.L == L) { *out = &es; return LPFNT_OK; }
    const int m = p->m, w = m - 1;
    std::vector<int32_t> l0;
    std::vector<uint8_t> al;
    const uint8_t *A = p->h_line_alpha.data();
    for (int64_t l = 0; l < p->N1; ++l) {
        const uint8_t *k = A + l * w + (w - L);
        if (l == 0 || std::memcmp(k, A + (l - 1) * w + (w - L), size_t(L)) != 0) {
            l0.push_back(int32_t(l));
            al.insert(al.end(), k, k + L);
        }
    }
    l0.push_back(int32_t(p->N1));
    lpfnt_plan::EvalSplit es;
    es.L = L;
    es.num_sub = int32_t(l0.size() - 1);
    int rc;
    if ((rc = upload(&es.d_l0, l0.data(), l0.size(), p))) return rc;
    if ((rc = upload(&es.d_alpha, al.data(), al.size(), p))) return rc;
    p->eval_splits.push_back(es);
    *out = &p->eval_splits.back();
    return LPFNT_OK;
}

// number of subtrees at level L without building the tables
int64_t eval_split_count(const lpfnt_plan *p, int L) {
    const int w = p->m - 1;
    const uint8_t *A = p->h_line_alpha.data();
    int64_t g = 0;
    for (int64_t l = 0; l < p->N1; ++l)
        if (l == 0 || std::memcmp(A + l * w + (w - L), A + (l - 1) * w + (w - L), size_t(L)) != 0) ++g;
    return g;
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
        static_assert(sizeof(FibEnt) == sizeof(int2), "FibEnt must match int2");
        static_assert(sizeof(SortedItem) == sizeof(int2), "SortedItem must match int2");
        const bool generic_only = D.groups.empty();  // fibres too long for the register kernels
        if (generic_only || hx.max_line > kMaxRegT || D.max_lines > kMaxRegT || N <= (int64_t(1) << 20)) {
            // the generic kernel walks the plain CSR
            if ((rc = upload(&d.thr_fib, D.thr_fib.data(), D.thr_fib.size(), p))) return fail(rc);
            if ((rc = upload(&d.thr_start, D.thr_start.data(), D.thr_start.size(), p))) return fail(rc);
            if ((rc = upload(&d.fib_ptr, D.fib_ptr.data(), D.fib_ptr.size(), p))) return fail(rc);
        }
        if ((rc = upload(&d.ent, reinterpret_cast<const int2 *>(D.ent.data()), D.ent.size(), p))) return fail(rc);
        if (!D.groups.empty()) {
            d.num_groups = int32_t(D.groups.size());
            if ((rc = upload(&d.groups, reinterpret_cast<const int2 *>(D.groups.data()), D.groups.size(), p))) return fail(rc);
            if ((rc = upload(&d.group_t, D.group_t.data(), D.group_t.size(), p))) return fail(rc);
            if ((rc = upload(&d.group_cta_cap, D.group_cta_cap.data(), D.group_cta_cap.size(), p))) return fail(rc);
        }
    }
    if ((rc = upload(&p->d_line_order, hx.line_order.data(), hx.line_order.size(), p))) return fail(rc);
    if ((rc = build_slabs(p, hx))) return fail(rc);
    if (hx.whole.ok || (m >= 2 && m <= 3 && N < 65536)) {
        // whole-function kernel: keep the index structure (its tables can be re-built with other options)
        p->hx = new HostIndex(hx);
        p->hx->whole = WholeHost();
        if (perm) {
            std::vector<uint16_t> p16(N);
            for (int64_t i = 0; i < N; ++i) p16[i] = uint16_t(perm[i]);
            if ((rc = upload(&p->d_perm16, p16.data(), p16.size(), p))) return fail(rc);
        }
        for (int which = 0; which < 2; ++which)
            if ((rc = upload_whole(p, which))) return fail(rc);
    }
    if (!hx.eval_block.empty()) {
        static_assert(sizeof(EvalBlock) == sizeof(int4), "EvalBlock must match int4");
        if ((rc = upload(&p->d_eval_line, reinterpret_cast<const int2 *>(hx.eval_line.data()), hx.eval_line.size(), p))) return fail(rc);
        if ((rc = upload(&p->d_eval_block, reinterpret_cast<const int4 *>(hx.eval_block.data()), hx.eval_block.size(), p))) return fail(rc);
        if (!hx.eval_tline.empty() && (rc = upload(&p->d_eval_tline, reinterpret_cast<const int2 *>(hx.eval_tline.data()), hx.eval_tline.size(), p))) return fail(rc);
        p->num_eval_blocks = int32_t(hx.eval_block.size());
        p->eval_padded = hx.eval_padded;
    }
    if (!hx.line_alpha.empty()) {
        // the eval kernel stages the alpha bytes in 4-byte words: pad the table so that its last word is complete
        std::vector<uint8_t> al(hx.line_alpha);
        al.resize(((al.size() + 3) & ~size_t(3)) + 4, 0);
        if ((rc = upload(&p->d_line_alpha, al.data(), al.size(), p))) return fail(rc);
        p->h_line_alpha = hx.line_alpha;
    }
    // the generic eval needs A on the device: keep a host copy when the fast path cannot run, or when the set is small
    // enough for the copy not to matter (force_generic in the tests)
    if (n + 1 > kMaxRegT || m > kEvalMaxM || n > 255 || N * m <= (int64_t(1) << 20)) p->hA.assign(A, A + N * m);
    if (nodes && (rc = upload(&p->d_nodes, nodes, size_t(n + 1), p))) return fail(rc);
    {
        cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking);
        if (e != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(LPFNT_ECUDA); }
    }
    *plan = p;
    return LPFNT_OK;
}

int lpfnt_plan_set_push(lpfnt_plan *p, const void *local_base, int64_t bytes, int n_peers, const void *const *peer_bases, const void *mc_base) {
    if (!p) { set_error("null plan"); return LPFNT_EINVAL; }
    if (n_peers == 0 || !local_base) { p->push = lpfnt_plan::Push(); return LPFNT_OK; }
    if (n_peers < 1 || n_peers > 8 || bytes <= 0 || !peer_bases) { set_error("set_push: need 1..8 peer mappings and a region size"); return LPFNT_EINVAL; }
    p->push.local = static_cast<const char *>(local_base);
    p->push.bytes = bytes;
    p->push.n = n_peers;
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_bases[r]) { set_error("set_push: null peer mapping"); return LPFNT_EINVAL; }
        p->push.peer[r] = static_cast<char *>(const_cast<void *>(peer_bases[r]));
    }
    p->push.mc = static_cast<char *>(const_cast<void *>(mc_base));
    return LPFNT_OK;
}

int lpfnt_push_rows_ex(int device, const void *src, int64_t bytes, int n_dst, void *const *dst, void *mc_dst, int ctas_per_sm, void *stream) {
    if (!src || bytes <= 0 || (bytes & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || n_dst < 0 || n_dst > 8 || (n_dst > 0 && !dst) || (n_dst == 0 && !mc_dst)) {
        set_error("push_rows: need a 16-byte aligned source, a multiple of 16 bytes and 1..8 destinations or a multicast mapping");
        return LPFNT_EINVAL;
    }
    CK(cudaSetDevice(device));
    PushArgs a{};
    a.src = static_cast<const uint4 *>(src);
    a.words = bytes / 16;
    a.mc = static_cast<uint4 *>(mc_dst);
    for (int r = 0; r < n_dst; ++r) {
        if (!dst[r] || (reinterpret_cast<uintptr_t>(dst[r]) & 15)) { set_error("push_rows: destinations must be 16-byte aligned"); return LPFNT_EINVAL; }
        if (dst[r] != src) a.dst[a.n_dst++] = static_cast<uint4 *>(dst[r]);  // the own mapping already holds the rows
    }
    if (!a.mc && a.n_dst == 0) return LPFNT_OK;
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) sms = 148;  // (cudaGetDeviceProperties takes milliseconds)
    // ctas_per_sm = 1 (lpfnt_push_rows): ONE small CTA per SM fits beside the persistent sweep CTA (4096 registers are left); two
    // would keep the next sweep CTA out of the SM until the copy has finished.  A copy that nothing follows (the last chunk of
    // a step) takes more: 8 CTAs per SM move 436 MB to seven peers in 0.60 instead of 0.69 ms.
    const int per_sm = std::max(1, std::min(ctas_per_sm, 16));
    const int grid = int(std::min<int64_t>((a.words + 127) / 128, int64_t(sms) * per_sm));
    cudaError_t e = launch_push_rows(a, grid, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { set_error(std::string("push_rows launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
    return LPFNT_OK;
}

int lpfnt_push_rows(int device, const void *src, int64_t bytes, int n_dst, void *const *dst, void *mc_dst, void *stream) {
    return lpfnt_push_rows_ex(device, src, bytes, n_dst, dst, mc_dst, 1, stream);
}

int lpfnt_push_rows_ce(int device, const void *src, int64_t bytes, int n_dst, void *const *dst, void *stream) {
    if (!src || bytes <= 0 || n_dst < 1 || n_dst > 8 || !dst) { set_error("push_rows_ce: bad arguments"); return LPFNT_EINVAL; }
    CK(cudaSetDevice(device));
    for (int r = 0; r < n_dst; ++r) {
        if (!dst[r]) { set_error("push_rows_ce: null destination"); return LPFNT_EINVAL; }
        if (dst[r] == src) continue;  // the own mapping already holds the rows
        CK(cudaMemcpyAsync(dst[r], src, size_t(bytes), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    }
    return LPFNT_OK;
}

int lpfnt_plan_destroy(lpfnt_plan *p) {
    if (!p) return LPFNT_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_eval_line); cudaFree(p->d_eval_tline); cudaFree(p->d_eval_block); cudaFree(p->d_cpad);
    for (auto &es : p->eval_splits) { cudaFree(es.d_l0); cudaFree(es.d_alpha); }
    free_whole(p, 0);
    free_whole(p, 1);
    free_wsplit(p);
    cudaFree(p->d_perm16);
    delete p->hx;
    cudaFree(p->d_timeline);
    cudaFree(p->slabs.d_desc); cudaFree(p->slabs.d_item0); cudaFree(p->slabs.d_item1);
    cudaFree(p->d_line_off); cudaFree(p->d_perm); cudaFree(p->d_line_order); cudaFree(p->d_line_alpha); cudaFree(p->d_A);
    cudaFree(p->d_nodes);
    for (OperatorCache *oc : p->ops) { cudaFree(oc->d_packed); delete oc; }
    for (int h = 0; h <= kMaxDim; ++h) {
        cudaFree(p->dd[h].thr_fib); cudaFree(p->dd[h].thr_start); cudaFree(p->dd[h].fib_ptr); cudaFree(p->dd[h].ent);
        cudaFree(p->dd[h].groups); cudaFree(p->dd[h].group_t); cudaFree(p->dd[h].group_cta_cap);
    }
    for (int s = 0; s < 6; ++s) cudaFree(p->ws[s]);
    for (int s = 0; s < 4; ++s) if (p->hstage[s]) cudaFreeHost(p->hstage[s]);
    for (int s = 0; s < 2; ++s) if (p->lane_done[s]) cudaEventDestroy(p->lane_done[s]);
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

constexpr int kTimelineCount = 4 * 2 * 24 * kWholeStamps;

int lpfnt_plan_set_option(lpfnt_plan *p, const char *name, int64_t value) {
    if (!p || !name) { set_error("bad arguments"); return LPFNT_EINVAL; }
    if (std::strcmp(name, "force_generic") == 0) { p->force_generic = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "trace") == 0) { p->trace = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "dx_descending") == 0) { p->dx_descending = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "eval_block") == 0) { p->use_eval_block = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "eval_split") == 0) { p->use_eval_split = value < 0 ? -1 : (value ? 1 : 0); return LPFNT_OK; }
    if (std::strcmp(name, "host_chunk_bytes") == 0) { p->host_chunk_bytes = value <= 0 ? 0 : std::max<int64_t>(value, 1 << 16); return LPFNT_OK; }
    if (std::strcmp(name, "host_threads") == 0) { p->host_threads = int(std::max<int64_t>(0, std::min<int64_t>(value, 64))); return LPFNT_OK; }
    if (std::strcmp(name, "use_cta_cap") == 0) { p->use_cta_cap = value ? 1 : 0; return LPFNT_OK; }
    if (std::strcmp(name, "timeline") == 0) {
        CK(cudaSetDevice(p->device));
        if (value && !p->d_timeline) {
            CK(cudaMalloc(reinterpret_cast<void **>(&p->d_timeline), kTimelineCount * sizeof(long long)));
            CK(cudaMemset(p->d_timeline, 0, kTimelineCount * sizeof(long long)));
        } else if (!value && p->d_timeline) { CK(cudaDeviceSynchronize()); cudaFree(p->d_timeline); p->d_timeline = nullptr; }
        return LPFNT_OK;
    }
    if (std::strcmp(name, "slab_ctas") == 0) { p->slab_ctas = int(std::max<int64_t>(1, std::min<int64_t>(value, 4))); return LPFNT_OK; }
    if (std::strcmp(name, "use_slab") == 0) {
        p->use_slab = value ? 1 : 0;
        if (value == 2 && p->slabs.n_slabs > 0) p->slabs.ok = true;  // 2: force the slab kernel even for sparsely filled slabs (tests)
        return LPFNT_OK;
    }
    if (std::strcmp(name, "use_whole") == 0) { p->use_whole = value < 0 ? -1 : (value ? 1 : 0); return LPFNT_OK; }
    if (std::strcmp(name, "whole_split_skew") == 0) { p->whole_split_skew = int(std::max<int64_t>(0, std::min<int64_t>(value, 1000000))); return LPFNT_OK; }
    if (std::strcmp(name, "whole_split") == 0) {
        p->use_whole_split = value < 0 ? -1 : (value ? 1 : 0);
        if (p->use_whole_split != 0 && !p->wsplit.h.ok) {  // the tables (16 MB of scratch at d = 3, n = 30) are built on first request
            CK(cudaSetDevice(p->device));
            CK(cudaDeviceSynchronize());
            return upload_wsplit(p);
        }
        return LPFNT_OK;
    }
    if (std::strcmp(name, "whole_split_c") == 0 || std::strcmp(name, "whole_split_order") == 0) {
        if (name[12] == 'c') p->whole_split_c = int(std::max<int64_t>(0, std::min<int64_t>(value, 16)));
        else p->whole_split_order = value < 0 ? kWholeDefaultOrder : int(value & 127);
        if (p->use_whole_split == 0) return LPFNT_OK;  // used when the split kernel is switched on
        CK(cudaSetDevice(p->device));
        CK(cudaDeviceSynchronize());
        return upload_wsplit(p);
    }
    if (std::strncmp(name, "whole_", 6) == 0) {
        // whole_<threads|fibres|order|rows>[_exact|_fma]: CTA shape of the whole-function kernel, for both forms or one
        std::string key(name + 6);
        int lo = 0, hi = 1;
        if (key.size() > 6 && key.compare(key.size() - 6, 6, "_exact") == 0) { hi = 0; key.resize(key.size() - 6); }
        else if (key.size() > 4 && key.compare(key.size() - 4, 4, "_fma") == 0) { lo = 1; key.resize(key.size() - 4); }
        if (key == "threads" || key == "fibres" || key == "order" || key == "rows") {
            for (int w = lo; w <= hi; ++w) {
                lpfnt_plan::WholeCfg &c = p->wcfg[w];
                if (key == "threads") c.threads = (value == 384 || value == 768) ? int(value) : -1;
                else if (key == "fibres") c.fibres = (value >= 1 && value <= 2) ? int(value) : -1;
                else if (key == "order") c.order = value < 0 ? kWholeDefaultOrder : int(value & 255);
                else c.rows = (value == 2 || value == 4) ? int(value) : -1;
            }
            if (key == "rows") return LPFNT_OK;
            CK(cudaSetDevice(p->device));
            CK(cudaDeviceSynchronize());  // the tables may be in use
            for (int w = lo; w <= hi; ++w) { int rc = upload_whole(p, w); if (rc) return rc; }
            return LPFNT_OK;
        }
    }
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
    CK(cudaMemcpy(out, p->d_timeline, sizeof(long long) * size_t(std::min(count, kTimelineCount)), cudaMemcpyDeviceToHost));
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
    spec.desc = p->dx_descending != 0;
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
                      (!p->chebyshev_eval || (p->use_eval_block && p->d_eval_block && p->eval_padded > 0));
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
    // few points: one thread per point cannot fill the GPU (the kernels below walk all N coefficients per thread) -- cut the
    // index trie into at least ~4 subtrees per SM and point tile, evaluate them side by side, combine in basis form
    bool split_done = false;
    if (fast && !p->chebyshev_eval && p->m >= 2 && p->use_eval_split != 0 && !p->h_line_alpha.empty()) {
        const int64_t tiles = (P + 255) / 256;
        // enough subtrees to fill the GPU, and subtrees short enough (<= ~512 coefficients) for the per-thread chains
        const int64_t want = std::max<int64_t>((int64_t(4) * p->num_sms + tiles - 1) / tiles, p->N / 512);
        if (p->use_eval_split == 1 || (tiles < 2 * p->num_sms && P <= 8192)) {
            int L = 0;
            int64_t G = 0;
            if (p->eval_split_counts.empty())
                for (int l = 1; l <= p->m - 1; ++l) p->eval_split_counts.push_back(eval_split_count(p, l));
            for (int l = 1; l <= p->m - 1; ++l) {
                const int64_t g = p->eval_split_counts[l - 1];
                if (g > 65536 && L > 0) break;
                L = l; G = g;
                if (g >= want) break;
            }
            if (L > 0 && G > 1) {
                const lpfnt_plan::EvalSplit *es = nullptr;
                int rc = eval_split_level(p, L, &es);
                if (rc) return rc;
                // partial sums: G doubles per point, at most 256 MB at a time
                const int64_t pc = std::max<int64_t>(1, std::min<int64_t>(P, (int64_t(256) << 20) / (8 * G)));
                rc = ensure_ws(p, 2, size_t(G) * size_t(pc) * sizeof(double));
                if (rc) return rc;
                for (int64_t q0 = 0; q0 < P; q0 += pc) {
                    const int64_t np = std::min(pc, P - q0);
                    EvalSubArgs a{dc, dp + q0 * p->m, p->ws[2], p->d_line_off, p->d_line_alpha, es->d_l0, es->d_alpha, dv + q0, np, es->num_sub, L, p->m, p->n + 1};
                    trace_begin(p, 1300, s);
                    cudaError_t e = launch_eval_split(a, p->nodes.data(), p->n + 1, s);
                    trace_end(p, s);
                    p->launches += 2;
                    if (e != cudaSuccess) { set_error(std::string("eval split launch: ") + cudaGetErrorString(e)); return LPFNT_ECUDA; }
                }
                split_done = true;
            }
        }
    }
    if (split_done) {
    } else if (fast) {
        const bool blocked = p->use_eval_block && p->d_eval_block && p->eval_padded > 0 && (p->m == 1 || p->d_line_alpha) && p->d_eval_tline;
        if (blocked && !p->d_cpad) {
            if (cudaMalloc(reinterpret_cast<void **>(&p->d_cpad), sizeof(double) * size_t(p->eval_padded + 4)) != cudaSuccess) { set_error("cudaMalloc(cpad) failed"); return LPFNT_ENOMEM; }
            p->table_bytes += int64_t(sizeof(double)) * (p->eval_padded + 4);
        }
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
                EvalBlockArgs a{p->d_cpad, dp + q0 * p->m, dv + q0, p->d_eval_block, p->d_eval_line, p->d_eval_tline, p->d_line_alpha, np, p->num_eval_blocks, p->m};
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
