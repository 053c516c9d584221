"""Row sharding of batched transforms across GPUs (one process per GPU).

The reference has no distributed code.  Batches of independent functions (rows of
the (B, N) array) and evaluation point sets shard trivially: every rank owns a contiguous block of
rows, runs the single-GPU path on it and -- only if the caller wants the full result everywhere --
a final all-gather collects the blocks (NCCL over NVLink on GPUs, gloo on CPU in the tests).
No collective sits on the transform path itself.
"""
from __future__ import annotations

from typing import Tuple

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def row_block(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous block; the first ``total % world`` ranks get one extra row."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("row_block: need world >= 1, 0 <= rank < world, total >= 0")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local, total: int, group=None):
    """All-gather row blocks produced with :func:`row_block` into the full (total, ...) tensor.

    Works for uneven blocks (pads to the largest block for the collective, trims afterwards)."""
    if dist is None or not dist.is_initialized():
        if local.shape[0] != total:
            raise ValueError("gather_rows without an initialised process group needs the full array")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, stop = row_block(total, rank, world)
    if local.shape[0] != stop - start:
        raise ValueError(f"rank {rank}: expected {stop - start} rows, got {local.shape[0]}")
    biggest = -(-total // world)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        a, b = row_block(total, r, world)
        pieces.append(out[r * biggest : r * biggest + (b - a)])
    return torch.cat(pieces, dim=0)


def plan_push_chunks(rows: int, rows_per_round: int, sweep_ms_per_round: float, copy_ms_per_row: float,
                     launch_ms: float = 0.008, max_chunks: int = 5):
    """Cut this rank's `rows` functions into chunks for a pipelined gather: chunk k is copied to the peers on a side stream while
    chunk k+1 is swept.  Chunks are whole rounds of the persistent sweep kernel (`rows_per_round` = one function per SM).  The
    cut minimises the modelled end of the last copy: a sweep of r rounds takes r * sweep_ms_per_round + launch_ms, a copy of n
    rows n * copy_ms_per_row + launch_ms, copies run one after another.  When the sweeps dominate the chunks shrink
    geometrically (only the last, small copy is exposed); when the copies dominate the first chunk is one round (the link
    starts early) and the rest is sent in few large pieces.  Returns [(lo, hi), ...]."""
    if rows <= 0:
        return []
    rounds = -(-rows // rows_per_round)
    best, best_cut = None, None

    def finish(parts):
        t_sweep = t_copy = 0.0
        done = 0
        for r in parts:
            n = min(rows, (done + r) * rows_per_round) - done * rows_per_round
            t_sweep += r * sweep_ms_per_round + launch_ms
            t_copy = max(t_copy, t_sweep) + n * copy_ms_per_row + launch_ms
            done += r
        return t_copy

    def rec(left, parts):
        nonlocal best, best_cut
        if left == 0:
            f = finish(parts)
            if best is None or f < best - 1e-12:
                best, best_cut = f, list(parts)
            return
        if len(parts) == max_chunks - 1:
            rec(0, parts + [left])
            return
        for r in range(1, left + 1):
            rec(left - r, parts + [r])

    rec(rounds, [])
    out, done = [], 0
    for r in best_cut:
        lo, hi = done * rows_per_round, min(rows, (done + r) * rows_per_round)
        if hi > lo:
            out.append((lo, hi))
        done += r
    return out


class PushGather:
    """All-gather of the row blocks of a sharded batch WITHOUT a collective: the full (total, N) result lives in a symmetric
    buffer (``torch.distributed._symmetric_memory``) that every rank maps -- also through the NVSwitch multicast mapping
    when the fabric offers one -- and the whole-function kernel stores every result element of this rank's rows into the same
    offset of every rank's copy (``lpfnt_plan_set_push``): one ``multimem.st`` per element, or one store per peer over NVLink.
    The ranks only synchronise afterwards (:meth:`barrier`).

        pg = PushGather(t, total_rows)           # collective: every rank calls it
        t.ifnt(coef_block, out=pg.local_rows())  # this rank's rows -> every rank's buffer
        pg.barrier(); full = pg.full             # (total, N) on every rank
    """

    def __init__(self, t, total: int, group=None, use_multicast: bool = True, fused: bool = True):
        """fused=True registers the buffer with the plan (the whole-function kernel stores into every rank's copy);
        fused=False leaves the kernels alone: write ``out=local_rows()`` as usual, then :meth:`push_rows` (a streaming copy
        kernel on a stream of your choice) sends row ranges to the other ranks."""
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _native as nat

        if dist is None or not dist.is_initialized():
            raise RuntimeError("PushGather needs an initialised torch.distributed process group")
        self.t, self.total = t, int(total)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PushGather: at most 8 ranks (one NVSwitch domain)")
        N = len(t)
        dev = torch.device("cuda", t.device)
        self.full = symm_mem.empty((self.total, N), dtype=torch.float64, device=dev)
        self.hdl = symm_mem.rendezvous(self.full, self.group)
        peers = [int(p) for p in self.hdl.buffer_ptrs]
        # the handle's pointers address the allocation; the tensor may sit at an offset inside it
        off = int(self.hdl.offset) if hasattr(self.hdl, "offset") else 0
        base = self.full.data_ptr()
        if peers[self.rank] + off != base:
            off = base - peers[self.rank]
        mc = int(self.hdl.multicast_ptr) if (use_multicast and self.hdl.has_multicast_support) else 0
        self.multicast = mc != 0
        arr = (C.c_void_p * self.world)(*[p + off for p in peers])
        nbytes = self.total * N * 8
        self._peer_bases = [p + off for p in peers]
        self._mc_base = mc + off if mc else 0
        self._row_bytes = N * 8
        self.fused = bool(fused)
        if self.fused:
            nat.check(nat.lib().lpfnt_plan_set_push(t._plan, C.c_void_p(base), nbytes, self.world, arr, C.c_void_p(mc + off if mc else 0)))
        self._a, self._b = row_block(self.total, self.rank, self.world)

    def local_rows(self):
        """This rank's rows of the full result (pass as ``out=``)."""
        return self.full[self._a:self._b]

    def push_rows(self, lo: int, hi: int, stream=None, engine: str = "sm", alone: bool = False):
        """Send the rows [lo, hi) of THIS rank's block (already written into its own copy) to every other rank's copy:
        engine="sm" a streaming copy kernel (multimem.st.v4 when the buffer has a multicast mapping; one small CTA per SM
        beside the sweep kernels, eight when ``alone`` says that no sweep follows), engine="ce" the copy engines (they need
        no SM)."""
        import ctypes as C

        from . import _native as nat

        if not (0 <= lo <= hi <= self._b - self._a):
            raise ValueError("push_rows: row range outside this rank's block")
        if hi == lo:
            return
        s = stream if stream is not None else torch.cuda.current_stream(self.full.device)
        first = (self._a + lo) * self._row_bytes
        nbytes = (hi - lo) * self._row_bytes
        dst = (C.c_void_p * self.world)(*[b + first for b in self._peer_bases])
        if engine == "ce":
            # one stream per peer: the copies to different peers run on different copy engines at the same time
            if not hasattr(self, "_peer_streams"):
                self._peer_streams = [torch.cuda.Stream(self.full.device) for _ in range(self.world)]
            src = C.c_void_p(self._peer_bases[self.rank] + first)
            for q in range(1, self.world):
                r = (self.rank + q) % self.world  # every rank starts with another peer
                ps = self._peer_streams[r]
                ps.wait_stream(s)
                one = (C.c_void_p * 1)(self._peer_bases[r] + first)
                nat.check(nat.lib().lpfnt_push_rows_ce(self.t.device, src, nbytes, 1, one, C.c_void_p(ps.cuda_stream)))
            for q in range(1, self.world):
                s.wait_stream(self._peer_streams[(self.rank + q) % self.world])
            return
        mc = C.c_void_p(self._mc_base + first) if self._mc_base else C.c_void_p(0)
        nat.check(nat.lib().lpfnt_push_rows_ex(self.t.device, C.c_void_p(self._peer_bases[self.rank] + first), nbytes,
                                               0 if self._mc_base else self.world, dst, mc, 8 if alone else 1,
                                               C.c_void_p(s.cuda_stream)))

    def barrier(self):
        """Every rank's stores have landed everywhere (device-side barrier on the current stream)."""
        self.hdl.barrier()

    def close(self):
        from . import _native as nat

        if self.fused:
            nat.check(nat.lib().lpfnt_plan_set_push(self.t._plan, None, 0, 0, None, None))


def sharded_apply(fn, full_rows, group=None, gather: bool = True):
    """Apply ``fn`` (e.g. ``t.fnt``) to this rank's block of ``full_rows`` and optionally all-gather."""
    if dist is None or not dist.is_initialized():
        return fn(full_rows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a, b = row_block(full_rows.shape[0], rank, world)
    local = fn(full_rows[a:b])
    return gather_rows(local, full_rows.shape[0], group) if gather else local


# ---------------------------------------------------------------------------------------------
# Sharded SINGLE transform: one coefficient vector too large (or too slow)
# for one GPU.  The sweeps along coordinates 0..m-2 only couple elements with the same last
# coordinate k = alpha_{m-1}, the sweep along coordinate m-1 only elements with the same
# (alpha_0..alpha_{m-2}).  So
#   * "slab" layout : rank r owns a run of planes k in [ks[r], ks[r+1]) -- a contiguous piece of the
#                     internal (colex) order; shifted by ks[r] it is itself a downward-closed set,
#   * "column" layout: rank r owns the complete fibres along coordinate m-1 of a run of base points
#                     b in [bs[r], bs[r+1]) (base = the plane k = 0), longest column first -- as the
#                     2-D set {(k, j) : k < height_j} again downward closed, columns = its lines,
#   * one all-to-all (NCCL over NVLink / NVSwitch) moves a vector from one layout to the other.
# fnt  = slab sweeps (0..m-2) -> all-to-all -> column sweep (m-1); result in column layout.
# ifnt = all-to-all -> slab sweeps -> all-to-all -> column sweep -> all-to-all; result in slab layout.
# The operation order per element is the single-GPU one, so the results are bit-identical to it.
# ---------------------------------------------------------------------------------------------
import numpy as np


class SlabColumnLayout:
    """Index logic of the two layouts and of the exchange between them (NumPy only, no GPU, no communication).

    Every rank builds the same object from the full multi-index array A (colex order)."""

    def __init__(self, A: np.ndarray, world: int):
        from .core import set as lpset

        A = np.ascontiguousarray(A, dtype=np.int64)
        N, m = A.shape
        if m < 2:
            raise ValueError("a 1-D transform has a single fibre and cannot be sharded")
        if world < 1:
            raise ValueError("world must be >= 1")
        self.N, self.m, self.world = N, m, world
        k = A[:, m - 1]
        K = int(k[-1]) + 1
        self.K = K
        plane = np.searchsorted(k, np.arange(K + 1))  # plane k = elements [plane[k], plane[k+1])
        # slab partition: runs of planes with about N / world elements each
        ks = [0]
        for r in range(1, world):
            ks.append(int(np.clip(np.searchsorted(plane, r * N / world), ks[-1], K)))
        ks.append(K)
        self.ks = np.asarray(ks)
        self.slab_elem = plane[self.ks]  # element range of every rank's slab
        # base index (position of (alpha_0..alpha_{m-2}) in plane 0) of every element
        base_set = np.ascontiguousarray(A[: plane[1], : m - 1])
        nb = len(base_set)
        base = np.empty(N, dtype=np.int64)
        base[: plane[1]] = np.arange(nb)
        for q in range(1, K):
            base[plane[q] : plane[q + 1]] = lpset.ordinal_embedding(np.ascontiguousarray(A[plane[q] : plane[q + 1], : m - 1]), base_set)
        height = np.bincount(base, minlength=nb)
        # column partition: runs of base points with about N / world elements each
        cum = np.concatenate(([0], np.cumsum(height)))
        bs = [0]
        for r in range(1, world):
            bs.append(int(np.clip(np.searchsorted(cum, r * N / world), bs[-1], nb)))
        bs.append(nb)
        self.bs = np.asarray(bs)
        owner = np.searchsorted(self.bs, base, side="right") - 1  # column owner of every element
        # column layout of every rank: its base points sorted by height (longest first, stable)
        col_pos = np.empty(N, dtype=np.int64)  # position of every element inside its owner's column layout
        self.col_heights = []
        for r in range(world):
            b0, b1 = self.bs[r], self.bs[r + 1]
            h = height[b0:b1]
            order = np.argsort(-h, kind="stable")
            hs = h[order]
            start_sorted = np.concatenate(([0], np.cumsum(hs)))[:-1]
            start = np.empty(b1 - b0, dtype=np.int64)
            start[order] = start_sorted
            self.col_heights.append(hs)
            sel = owner == r
            col_pos[sel] = start[base[sel] - b0] + k[sel]
        self.col_elems = np.asarray([int(hh.sum()) for hh in self.col_heights])
        # exchange lists, slab -> column: sender s packs its elements by destination (stable in element order)
        self.send_idx, self.send_counts, self.recv_pos, self.recv_counts = [], [], [], []
        for s in range(world):
            e0, e1 = self.slab_elem[s], self.slab_elem[s + 1]
            dest = owner[e0:e1]
            order = np.argsort(dest, kind="stable")
            self.send_idx.append(order)
            self.send_counts.append(np.bincount(dest, minlength=world))
        for r in range(world):
            pos, cnt = [], []
            for s in range(world):
                e0, e1 = self.slab_elem[s], self.slab_elem[s + 1]
                sel = np.nonzero(owner[e0:e1] == r)[0] + e0
                pos.append(col_pos[sel])
                cnt.append(len(sel))
            self.recv_pos.append(np.concatenate(pos) if pos else np.zeros(0, dtype=np.int64))
            self.recv_counts.append(np.asarray(cnt))
        self._A = A
        self._plane = plane

    def slab_set(self, r: int) -> np.ndarray:
        """Multi-indices of rank r's slab, last coordinate shifted to start at 0 (downward closed, colex order)."""
        e0, e1 = self.slab_elem[r], self.slab_elem[r + 1]
        S = self._A[e0:e1].copy()
        if len(S):
            S[:, self.m - 1] -= self.ks[r]
        return S

    def column_set(self, r: int) -> np.ndarray:
        """The 2-D set {(k, j) : k < height_j} of rank r's columns (coordinate 0 = k runs fastest)."""
        hs = self.col_heights[r]
        j = np.repeat(np.arange(len(hs)), hs)
        kk = np.arange(int(hs.sum())) - np.repeat(np.concatenate(([0], np.cumsum(hs)))[:-1], hs)
        return np.ascontiguousarray(np.stack([kk, j], axis=1).astype(np.int64))

    # reference permutations for tests: full internal-order vector <-> concatenated layouts
    def full_to_columns(self, full: np.ndarray, r: int) -> np.ndarray:
        out = np.empty(self.col_elems[r], dtype=full.dtype)
        off = 0
        for s in range(self.world):
            e0, e1 = self.slab_elem[s], self.slab_elem[s + 1]
            loc = full[e0:e1][self.send_idx[s]]
            before = int(self.send_counts[s][:r].sum())
            cnt = int(self.send_counts[s][r])
            out[self.recv_pos[r][off : off + cnt]] = loc[before : before + cnt]
            off += cnt
        return out


class ShardedTransform:
    """One transform sharded over the ranks of a process group (see the block comment above).

    ``t`` is a constructed :class:`lpfun_b200.Transform` (used for its index set and 1-D matrices; build it with
    ``colex_order=False`` -- the sharded vectors live in the internal order).  ``backend="cuda"`` runs the sweeps
    with the library's kernels on this rank's GPU and exchanges with NCCL; a callable
    ``backend(A_local, packed, kind, solve, fma, dims, x) -> y`` (NumPy) is used by the CPU tests with gloo.
    """

    def __init__(self, t, group=None, backend="cuda"):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("ShardedTransform needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.t = t
        self.m, self.n = t.spatial_dimension, t.polynomial_degree
        self.layout = SlabColumnLayout(t.multi_index_set, self.world)
        L = self.layout
        self.backend = backend
        self._slab_A = L.slab_set(self.rank)
        self._col_A = L.column_set(self.rank)
        self._slab_plan = self._col_plan = None
        if backend == "cuda":
            import ctypes as C

            from . import _native as nat

            self._nat, self._C = nat, C
            self.device = torch.device("cuda", t.device)
            self._slab_plan = self._make_plan(self._slab_A, self.m)
            self._col_plan = self._make_plan(self._col_A, 2)
            dev = self.device
        else:
            dev = torch.device("cpu")
        self._send_idx = torch.as_tensor(L.send_idx[self.rank], device=dev)
        self._recv_pos = torch.as_tensor(L.recv_pos[self.rank], device=dev)
        self._send_counts = [int(c) for c in L.send_counts[self.rank]]
        self._recv_counts = [int(c) for c in L.recv_counts[self.rank]]
        self.exchanges = 0

    # -- plans -------------------------------------------------------------------------------
    def _make_plan(self, A, m):
        if len(A) == 0:
            return None
        nat, C = self._nat, self._C
        plan = C.c_void_p()
        A = np.ascontiguousarray(A, dtype=np.int64)
        nat.check(nat.lib().lpfnt_plan_create(m, self.n, len(A), A.ctypes.data, None, None, None, None, None, None, self.t.device, C.byref(plan)))
        return plan

    def close(self):
        if self.backend == "cuda":
            for p in (self._slab_plan, self._col_plan):
                if p:
                    self._nat.lib().lpfnt_plan_destroy(p)
            self._slab_plan = self._col_plan = None

    # -- local sweeps ------------------------------------------------------------------------
    def _sweep(self, which, spec, x):
        packed, kind, solve, fma = spec
        if x.numel() == 0:
            return x
        if which == "slab":
            A, plan, dims = self._slab_A, self._slab_plan, list(range(self.m - 1))
        else:
            A, plan, dims = self._col_A, self._col_plan, [0]
        if self.backend == "cuda":
            nat = self._nat
            mask = 0
            for h in dims:
                mask |= 1 << h
            x = x.contiguous()
            out = torch.empty_like(x)
            stream = self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            nat.check(nat.lib().lpfnt_sweep(plan, packed.ctypes.data, kind, int(solve), int(fma), mask, x.data_ptr(), out.data_ptr(), 1, 1, stream))
            return out
        return torch.as_tensor(self.backend(A, packed, kind, solve, fma, dims, x.numpy()))

    # -- exchange ----------------------------------------------------------------------------
    def _to_columns(self, slab):
        send = slab[self._send_idx]
        recv = torch.empty(sum(self._recv_counts), dtype=slab.dtype, device=slab.device)
        dist.all_to_all_single(recv, send, self._recv_counts, self._send_counts, group=self.group)
        self.exchanges += 1
        cols = torch.empty_like(recv)
        cols[self._recv_pos] = recv
        return cols

    def _to_slab(self, cols):
        send = cols[self._recv_pos]
        recv = torch.empty(sum(self._send_counts), dtype=cols.dtype, device=cols.device)
        dist.all_to_all_single(recv, send, self._send_counts, self._recv_counts, group=self.group)
        self.exchanges += 1
        slab = torch.empty_like(recv)
        slab[self._send_idx] = recv
        return slab

    # -- operator sequences (transform.py:455-551, 578-632) --------------------------------------
    def _specs(self, op):
        from . import _native as nat

        t = self.t
        LOWER, UPPER = nat.LPFNT_LOWER, nat.LPFNT_UPPER
        if op == "fnt":
            if t._Vx_ut is not None:
                if t._inv_Vx_lt is None:
                    raise NotImplementedError("sharded fnt of an LU-factorised basis needs precomputation=True")
                return [(t._inv_Vx_lt, LOWER, False, False), (t._inv_Vx_ut, UPPER, False, False)]
            if t._inv_Vx_lt is not None:
                return [(t._inv_Vx_lt, LOWER, False, False)]
            return [(t._Vx_lt, LOWER, True, False)]
        fma = not t._exact_ifnt
        if t._Vx_ut is not None:
            return [(t._Vx_ut, UPPER, False, fma), (t._Vx_lt, LOWER, False, fma)]
        return [(t._Vx_lt, LOWER, False, fma)]

    def fnt(self, values_slab):
        """Function values of this rank's slab (internal order) -> coefficients of this rank's columns."""
        x, cols = values_slab, None
        for i, spec in enumerate(self._specs("fnt")):
            if i > 0:
                x = self._to_slab(cols)
            cols = self._sweep("cols", spec, self._to_columns(self._sweep("slab", spec, x)))
        return cols

    def ifnt(self, coeffs_cols):
        """Coefficients of this rank's columns -> function values of this rank's slab (internal order)."""
        cols = coeffs_cols
        for spec in self._specs("ifnt"):
            x = self._to_slab(cols)
            cols = self._sweep("cols", spec, self._to_columns(self._sweep("slab", spec, x)))
        return self._to_slab(cols)

    # -- helpers for callers that hold the full vector ---------------------------------------
    def slab_of(self, full):
        e0, e1 = self.layout.slab_elem[self.rank], self.layout.slab_elem[self.rank + 1]
        return full[e0:e1]

    def columns_of(self, full):
        """This rank's column layout of a full internal-order vector (host-side reference permutation)."""
        return self.layout.full_to_columns(np.asarray(full), self.rank)
