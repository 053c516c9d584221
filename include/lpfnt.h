/*
 * lpfnt.h -- C ABI of liblpfnt.so: the B200 (sm_100a) Fast Newton Transform path.
 *
 * This is the drop-in boundary for lpFun's data-parallel hot path.  The
 * reference (pure Python + Numba) has no FFI; the boundary sits exactly where
 * its `Transform` methods call the njit dispatchers
 *   src/lpfun/transform.py:462,477,582,679,765,844
 *     -> itransform / transform / dtransform  (src/lpfun/core/molecules.py:127,37,222)
 *     -> newton2point                         (src/lpfun/utils.py:132)
 * and where its constructor builds the index structure
 *   src/lpfun/transform.py:209-232 -> lp_set / lp_tube / entropy (src/lpfun/core/set.py:58,96,167).
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success, a negative
 *    LPFNT_E* code on failure (message: lpfnt_last_error(), thread local),
 *    except the *_count/_size helpers which return a size (or a negative code).
 *  - vectors are float64, laid out as (batch, N) row-major; `batch` = 1 is the
 *    reference's single-vector contract.
 *  - `on_device` = 0: in/out are HOST pointers, the call does H2D, compute, D2H
 *    and returns after the result is in `out` (NumPy drop-in path).
 *    `on_device` = 1: in/out are DEVICE pointers on the plan's device; the call
 *    only enqueues work on `stream` (a cudaStream_t passed as void*, NULL = the
 *    legacy default stream) and returns immediately.
 *  - the caller owns in/out; a plan owns its device tables and workspaces.  A plan
 *    may be used from one host thread at a time.
 *  - in == out (in-place) is allowed for every op.
 */
#ifndef LPFNT_H
#define LPFNT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpfnt_plan lpfnt_plan;

enum {
    LPFNT_OK = 0,
    LPFNT_EINVAL = -1,   /* bad argument (maps to ValueError in the Python front-end) */
    LPFNT_ECUDA = -2,    /* CUDA runtime error, text in lpfnt_last_error() */
    LPFNT_ENOMEM = -3,
    LPFNT_EUNSUPPORTED = -4
};

/* triangular operator kinds (replace the `mode` strings of molecules.py) */
enum {
    LPFNT_LOWER = 0,     /* packed lower triangle, row k at k(k+1)/2          (get_rmo, utils.py:295) */
    LPFNT_UPPER = 1      /* packed upper triangle, row k at k*(n+1)-k(k-1)/2  (transform.py:326)     */
};

/* arithmetic contract of a sweep */
enum {
    LPFNT_ARITH_EXACT = 0, /* reference order, multiply and add rounded separately: bit-identical */
    LPFNT_ARITH_FMA = 1    /* same order, fused multiply-add (<= 1 ulp per term from EXACT)        */
};

const char *lpfnt_version(void);
const char *lpfnt_last_error(void);
int lpfnt_device_count(int *count);

/* ---- host-side index structure (no GPU needed) --------------------------------------- */

/* |A| for the l^p lower set {a in N^m : sum a_j^p <= n^p} (p = INFINITY: full grid).
 * Replaces lp_set, src/lpfun/core/set.py:12-74 (same floating-point membership rule). */
int64_t lpfnt_lp_set_count(int m, int n, double p);
/* Fill A (N x m, int64, row-major) in colexicographic order; N must equal the count. */
int lpfnt_lp_set_fill(int m, int n, double p, int64_t *A, int64_t N);
/* Line lengths T along coordinate 0 (lp_tube, set.py:80-107). Returns N_1; T may be NULL. */
int64_t lpfnt_lp_tube(const int64_t *A, int64_t N, int m, int64_t *T);
/* Level sizes [N_0, N_1, ..., n+1, 1] of the index trie (entropy, set.py:167-182).
 * e_T must hold m+1 entries; returns the number written. */
int lpfnt_entropy(const int64_t *A, int64_t N, int m, int64_t *e_T);
/* Positions of the lower set A_small inside A_big, both colex ordered
 * (ordinal_embedding, set.py:208-239). phi holds N_small entries. */
int lpfnt_ordinal_embedding(const int64_t *A_small, int64_t N_small, const int64_t *A_big,
                            int64_t N_big, int m, int64_t *phi);


/* ---- derived index tables (host only; what a plan uploads) ----------------------------
 * Exposed so that the table layout can be inspected and tested without a GPU.
 *   what: LPFNT_TAB_*; h: coordinate (1..m-1) for the per-coordinate tables, ignored otherwise.
 * lpfnt_index_size returns the number of ELEMENTS of the table, lpfnt_index_copy copies it
 * (int32 for all tables except LINE_ALPHA = uint8, LEVEL_SIZE = int64, FIB_ENT = int32 pairs
 * {line element offset, line length}). */
typedef struct lpfnt_index lpfnt_index;
enum {
    LPFNT_TAB_LINE_OFF = 0,   /* N_1 + 1 : element offset of every line (cs_T)                 */
    LPFNT_TAB_FIB_PTR = 1,    /* F_h + 1 : CSR offsets of the fibres of coordinate h           */
    LPFNT_TAB_FIB_ENT = 2,    /* 2 * N_1 : per fibre, its lines in ascending a_h {off, len}    */
    LPFNT_TAB_THR_START = 3,  /* F_h + 1 : first thread (element-fibre) of each fibre          */
    LPFNT_TAB_THR_FIB = 4,    /* #threads: thread -> fibre                                     */
    LPFNT_TAB_LINE_ALPHA = 5, /* N_1 * (m-1) uint8 : (a_1..a_{m-1}) of every line (eval)       */
    LPFNT_TAB_LEVEL_SIZE = 6, /* m + 1 int64 : N_0, N_1, ..., 1                                */
    LPFNT_TAB_GROUPS = 7,       /* 2 * #groups int32: runs of <= 8 consecutive lanes of a fibre, tallest first:
                                   {first ent entry, lane0 | height << 8 | lanes << 16}               */
    LPFNT_TAB_LINE_ORDER = 8,   /* N_1 int32: lines ordered by length inside every chunk of 128 lines  */
    LPFNT_TAB_GROUP_T = 9,      /* 8 * #groups uint8: height of every lane of every lane group         */
    LPFNT_TAB_WHOLE_ITEMS = 10, /* uint32: item lists of the whole-function kernel, all parts and coordinates back to back in
                                   thread order: bits 0..15 line offset (coordinate 0) / first entry in WHOLE_SEO,
                                   bits 16..21 length, bits 22..31 lane a_0                                      */
    LPFNT_TAB_WHOLE_SEO = 11,   /* uint16: per coordinate h >= 1 the line offsets in fibre order                */
    LPFNT_TAB_LINE_LAST = 12,   /* N_1 uint8: a_{m-1} of every line                                    */
    LPFNT_TAB_WSPLIT_ITEMS = 13, /* uint32: item lists of the split whole-function kernel (two parts x three coordinates)  */
    LPFNT_TAB_WSPLIT_SEO = 14    /* uint16: its row offsets (inside the part)                                                */
};
int lpfnt_index_create(const int64_t *A, int64_t N, int m, int n, lpfnt_index **index);
int lpfnt_index_destroy(lpfnt_index *index);
/* Item tables of the whole-function kernel (m = 2, 3; N < 65536).  lpfnt_index_whole_build re-builds them for CTAs of
 * `threads` threads (384 or 768), `fibres` (1 item per thread and round; 2: two of equal length class, folded side by
 * side) and an item order (2 bits per coordinate, 0 strict length order,
 * 1 / 2 classes of 8 / 16 rows, 3 natural; + 64: chunks of 32 items dealt to the warps by a model of the shared FP64 pipe
 * instead of the snake order over the four SM sub-partitions; + 128: the chunks of coordinate 0 go to the warps in reverse, so
 * that the warps leaving the previous function's last sweep first start on the long lines; < 0 = default).  lpfnt_index_whole_meta fills 10 int32: ok, threads,
 * longest item, most items of a list, {list offset, padded list length} per coordinate; returns the number written. */
int lpfnt_index_whole_build(lpfnt_index *index, int threads, int order, int fibres);
int lpfnt_index_whole_meta(const lpfnt_index *index, int32_t *meta);
/* Split whole-function kernel (m = 3): the function as two units, planes a_2 < C and a_2 >= C (C <= 0: balanced), so that two
 * CTAs fit on an SM.  Meta: 83 int32 = ok, C, {e0, ne} x 2 parts, {list offset, padded length} x 2 parts x 3 coordinates, number of
 * columns of part Q, cb[32], rq[32] (column (a_0, a_1) of Q = cb[a_1] + a_0 for a_0 < rq[a_1]).  Item word: bits 0..15 line offset
 * inside the part / first entry in WSPLIT_SEO, 16..21 length, 22..26 lane a_0, 27..31 plane a_2 (coordinate 1) / row a_1 (coordinate 2). */
int lpfnt_index_wsplit_build(lpfnt_index *index, int C, int order);
int lpfnt_index_wsplit_meta(const lpfnt_index *index, int32_t *meta);
int64_t lpfnt_index_size(const lpfnt_index *index, int what, int h);
int lpfnt_index_copy(const lpfnt_index *index, int what, int h, void *dst);

/* ---- pinned host buffers for the host-pointer path ----------------------------------- */
int lpfnt_host_alloc(void **ptr, int64_t bytes);
int lpfnt_host_free(void *ptr);

/* ---- plan ------------------------------------------------------------------------------
 * Replaces the state the reference keeps on the Transform object
 * (src/lpfun/transform.py:207-336): A, T, cs_T, V_2, cs_V_2, e_T, the colex
 * permutation and the packed 1-D matrices.  All arrays are host pointers and are
 * copied; any of the matrix pointers may be NULL if the corresponding op is unused.
 *   A      (N, m) int64 multi-indices, colex order
 *   perm   (N) int64 colex permutation P (fnt gathers in[P[i]], ifnt scatters out[P[i]]) or NULL
 *   nodes  (n+1) Leja-ordered nodes
 *   Vx_lt / inv_Vx_lt     packed lower V and inv(V)   ((n+1)(n+2)/2 doubles each)
 *   Dx_ut[k] / DxT_lt[k]  packed upper D^(k+1) / packed lower (D^(k+1))^T, k = 0..2
 */
int lpfnt_plan_create(int m, int n, int64_t N, const int64_t *A, const int64_t *perm,
                      const double *nodes, const double *Vx_lt, const double *inv_Vx_lt,
                      const double *const *Dx_ut, const double *const *DxT_lt, int device,
                      lpfnt_plan **plan);
/* Bases whose Vandermonde matrix V is not triangular (basis="chebyshev"): V = L U without pivoting
 * (get_lu, src/lpfun/utils.py:315-328).  Pass L as Vx_lt / inv(L) as inv_Vx_lt to lpfnt_plan_create and the
 * packed UPPER factors here (transform.py:292-302).  fnt then runs the lower sweep with inv(L) followed by the
 * upper sweep with inv(U) (transform.py:491-519), ifnt the upper sweep with U followed by the lower sweep
 * with L (:596-624).  chebyshev_eval != 0 makes eval use the T_k recurrence (chebyshev2point, utils.py:165). */
int lpfnt_plan_set_upper_factors(lpfnt_plan *plan, const double *Vx_ut, const double *inv_Vx_ut, int chebyshev_eval);
int lpfnt_plan_destroy(lpfnt_plan *plan);
/* Fused all-gather (multi-GPU, one process per GPU): `local_base` .. + bytes is THIS rank's mapping of a symmetric buffer whose
 * peer mappings (the own one included, n_peers <= 8) and NVSwitch multicast mapping (or NULL) are given.  A batched device-pointer
 * fnt / ifnt whose `out` lies inside the region and that runs on the whole-function kernel then stores every result element into
 * the same offset of EVERY rank's buffer: one multimem.st through the multicast mapping, else one store per peer mapping over
 * NVLink -- the gather of the row blocks is done by the last sweep's stores, no collective follows (the ranks only synchronise).
 * n_peers = 0 clears the region.  The reference has no distributed code; this serves BASELINE config 4 (batch sharded by rows). */
/* The same gather as a streaming copy kernel: `bytes` at `src` (this rank's rows, already in its own copy of the symmetric buffer)
 * are stored into the same place of every other rank's copy -- dst[r] are the peer mappings of `src` (the own one is skipped), or
 * mc_dst its multicast mapping (one multimem.st.v4 per 16 bytes reaches every GPU).  Runs on `stream`, so it overlaps the sweeps of
 * the next chunk on another stream. */
int lpfnt_push_rows(int device, const void *src, int64_t bytes, int n_dst, void *const *dst, void *mc_dst, void *stream);
/* ... with the number of copy CTAs (128 threads) per SM: 1 (what lpfnt_push_rows uses) runs beside the sweep kernels of the next
 * chunk, more (up to 16) is faster for a copy that no sweep follows (the last chunk of a step) */
int lpfnt_push_rows_ex(int device, const void *src, int64_t bytes, int n_dst, void *const *dst, void *mc_dst, int ctas_per_sm, void *stream);
/* ... and by the copy engines (one cudaMemcpyAsync per peer mapping): no SM is taken from the persistent sweep kernels, so the
 * transfer really runs behind them */
int lpfnt_push_rows_ce(int device, const void *src, int64_t bytes, int n_dst, void *const *dst, void *stream);
int lpfnt_plan_set_push(lpfnt_plan *plan, const void *local_base, int64_t bytes, int n_peers, const void *const *peer_bases, const void *mc_base);
/* bytes of device memory held by the plan (tables + workspaces) */
int64_t lpfnt_plan_device_bytes(const lpfnt_plan *plan);
/* number of kernel launches enqueued by this plan since creation */
int64_t lpfnt_plan_launch_count(const lpfnt_plan *plan);
/* options: "force_generic" (1 = route every op through the unbounded-n fallback kernels),
 *          "trace" (1 = record a CUDA-event pair around every kernel launch),
 *          "dx_descending" (1 = lpfnt_dx folds every row from its last term down, the order of the reference on
 *          p = inf sets with m > 1, src/lpfun/core/molecules.py:290-294; Transform sets it),
 *          "use_whole" (whole function in shared memory, one pass over HBM per transform, m = 2, 3:
 *          -1 = automatic (default: batches of >= 8 functions), 0 = off, 1 = always),
 *          "whole_threads" / "whole_fibres" / "whole_order" (CTA shape of the whole-function kernel: 384 or 768 threads;
 *          1 item per thread, 2 = two of equal length class folded side by side; item order, see
 *          lpfnt_index_whole_build; -1 = automatic; the suffixes "_exact" / "_fma"
 *          address the tables of the exact / the fused-multiply-add forms only),
 *          "whole_split" (m = 3: every function as two units -- planes a_2 < C and a_2 >= C -- so that two CTAs fit on an SM;
 *          0 = off (default: measured slower), 1 = on, -1 = automatic), "whole_split_c" (the split plane, 0 = balanced),
 *          "whole_split_order", "whole_split_skew" (cycles the second CTA of an SM starts late),
 *          "use_slab" (coordinates 0 and 1 in one pass over slabs of planes, any m >= 2: 1 = when the slabs
 *          fill the CTA (default), 0 = off, 2 = always), "use_cta_cap" (k_sweep_groups: per-CTA row capacity),
 *          "eval_block" (0 = non-staged eval kernel), "timeline" (development: clock64 stamps of the
 *          whole-function kernel), "host_chunk_bytes" (host-pointer calls: bytes per pipelined chunk, 0 = automatic),
 *          "host_threads" (host-pointer calls on PAGEABLE memory: threads that copy the chunks to / from the plan's pinned
 *          staging buffers; 0 = automatic, 1 = single-threaded), "eval_split" (eval of few points: -1 automatic, 0 off, 1 on).
 * A plan owns one set of workspaces: use it from one host thread and one stream at a time. */
int lpfnt_plan_set_option(lpfnt_plan *plan, const char *name, int64_t value);
/* Drain the trace: per recorded launch a label (100 * kernel kind + coordinate; kinds: 0 line
 * sweep, 1 fibre sweep, 2 generic sweep, 3 eval (unstaged), 8 permutation, 11/12 eval (pad / block), 14 whole-function kernel, 16 slab kernel) and its device time in ms.
 * Synchronises on the recorded events.  *count receives the number of records seen. */
/* Coarse <-> fine coefficient transfer through the index map of lpfnt_ordinal_embedding / Transform.embed
 * (src/lpfun/transform.py:850-899; README "adaptivity": fine[phi] = coarse).  restrict_ = 0: prolongation,
 * out (batch, n_big) = 0 except out[b][phi[e]] = in[b][e]; restrict_ = 1: out (batch, n_small), out[b][e] = in[b][phi[e]].
 * phi is a host array of n_small int64.  Synchronises `stream` before returning. */
int lpfnt_embed_apply(int device, const int64_t *phi, int64_t n_small, int64_t n_big, const double *in, double *out, int64_t batch,
                      int restrict_, int on_device, void *stream);
/* development aid: clock64 stamps of CTAs 0 and 1 of the whole-function kernel, [4 functions][2 CTAs][warps][16 stamps]
 * (enable with option "timeline" = 1; see tools/timeline.py) */
int lpfnt_plan_timeline_read(lpfnt_plan *plan, long long *out, int count);
int lpfnt_plan_trace_read(lpfnt_plan *plan, int max_records, int *labels, float *ms, int *count);

/* ---- operator level: what the njit dispatchers compute ---------------------------------
 * lpfnt_itransform: triangular MAT-VEC along every coordinate h = 0..m-1
 *                   (itransform, molecules.py:127-219; atoms.py itransform_{lt,ut}_*).
 * lpfnt_transform : triangular SOLVE (forward/backward substitution) along every coordinate
 *                   (transform, molecules.py:37-124; atoms.py transform_{lt,ut}_*).
 * lpfnt_dtransform: triangular mat-vec along coordinate i only
 *                   (dtransform, molecules.py:222-309; atoms.py dtransform_{lt,ut}_md).
 * `packed` is a HOST pointer to the packed (n+1)x(n+1) triangle; `kind` LPFNT_LOWER/UPPER.
 * gather_perm / scatter_perm: apply the plan's colex permutation on the way in / out
 * (apply_permutation, src/lpfun/core/utils.py:6-20).
 */
int lpfnt_itransform(lpfnt_plan *plan, const double *packed, int kind, int arith,
                     const double *in, double *out, int64_t batch, int gather_perm,
                     int scatter_perm, int on_device, void *stream);
/* The same triangular operator (mat-vec, or solve when solve != 0) along the coordinates selected by dims_mask
 * (bit h = coordinate h), in ascending order, in internal (A) order without permutation.  This is the building
 * block of the sharded single transform (lpfun_b200/sharding.py): coordinates 0..m-2 on a slab of planes,
 * the last coordinate on re-distributed columns. */
int lpfnt_sweep(lpfnt_plan *plan, const double *packed, int kind, int solve, int arith, uint32_t dims_mask,
                const double *in, double *out, int64_t batch, int on_device, void *stream);
int lpfnt_transform(lpfnt_plan *plan, const double *packed, int kind, const double *in,
                    double *out, int64_t batch, int gather_perm, int scatter_perm,
                    int on_device, void *stream);
int lpfnt_dtransform(lpfnt_plan *plan, const double *packed, int kind, int arith, int i,
                     const double *in, double *out, int64_t batch, int on_device, void *stream);
/* newton2point (src/lpfun/utils.py:132-162): values[l] = sum_r c[r] prod_j N_{A[r,j]}(points[l,j]).
 * coeffs (N) and points (P, m) / values (P) follow on_device. */
int lpfnt_newton2point(lpfnt_plan *plan, const double *coeffs, const double *points, int64_t P,
                       double *values, int on_device, void *stream);

/* ---- method level: Transform.fnt / ifnt / dx / dxT / eval with the plan's matrices ------
 * fnt : gather by P, lower mat-vec with inv(V) in the reference's exact order
 *       (or the forward-substitution solve with V when the plan has no inv(V))   transform.py:430-490
 * ifnt: lower mat-vec with V (arith selectable), scatter by P                      transform.py:553-632
 * dx  : upper mat-vec with D^k along coordinate i                                 transform.py:634-718
 * dxT : lower mat-vec with (D^k)^T along coordinate i                             transform.py:720-804
 * eval: lpfnt_newton2point with the plan's nodes                                  transform.py:806-848
 */
int lpfnt_fnt(lpfnt_plan *plan, const double *in, double *out, int64_t batch, int on_device, void *stream);
int lpfnt_ifnt(lpfnt_plan *plan, const double *in, double *out, int64_t batch, int arith, int on_device, void *stream);
int lpfnt_dx(lpfnt_plan *plan, const double *in, double *out, int i, int k, int64_t batch, int on_device, void *stream);
int lpfnt_dxT(lpfnt_plan *plan, const double *in, double *out, int i, int k, int64_t batch, int on_device, void *stream);
int lpfnt_eval(lpfnt_plan *plan, const double *coeffs, const double *points, int64_t P, double *values, int on_device, void *stream);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* Time `iters` launches of a dependent-DFMA micro-kernel on `device`; returns FP64 FMA/s
 * in *fma_per_s (the FP64-pipe denominator used for eval's roofline). */
int lpfnt_measure_fp64(int device, int iters, double *fma_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LPFNT_H */
