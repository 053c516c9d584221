"""NumPy emulation of the GPU work decomposition (one thread = one element-fibre), driven by the
REAL device tables that liblpfnt.so derives on the host.  Used by CPU tests to validate the
tables and the per-thread arithmetic order without a GPU.  Elementwise NumPy multiply/add round
separately, i.e. exactly like the EXACT kernels (__dmul_rn / __dadd_rn)."""
import numpy as np

from lpfun_b200 import _native as nat
from lpfun_b200 import utils as U


def thread_index_matrix(tabs: nat.IndexTables, h: int, n1: int):
    """(num_threads, n1) element positions of every element-fibre along h, -1 padded."""
    off = tabs.table(nat.TAB_LINE_OFF)
    if h == 0:
        nl = len(off) - 1
        lens = np.diff(off)
        idx = off[:-1, None] + np.arange(n1)[None, :]
        idx = np.where(np.arange(n1)[None, :] < lens[:, None], idx, -1)
        return idx
    ptr = tabs.table(nat.TAB_FIB_PTR, h)
    ent = tabs.table(nat.TAB_FIB_ENT, h)
    tstart = tabs.table(nat.TAB_THR_START, h)
    tfib = tabs.table(nat.TAB_THR_FIB, h)
    nthr = len(tfib)
    lane = np.arange(nthr) - tstart[tfib]
    idx = np.full((nthr, n1), -1, dtype=np.int64)
    alive = np.ones(nthr, dtype=bool)
    for k in range(n1):
        p = ptr[tfib] + k
        ok = alive & (p < ptr[tfib + 1])
        e = ent[np.minimum(p, len(ent) - 1)]
        ok &= e[:, 1] > lane
        idx[ok, k] = e[ok, 0] + lane[ok]
        alive = ok
    return idx


def apply_tri(X, M, mode):
    """X: (threads, n1) padded with zeros; M: dense (n1, n1).  Returns the swept fibres."""
    n1 = X.shape[1]
    Y = np.zeros_like(X)
    if mode == "lower":
        for k in range(n1):
            acc = np.zeros(len(X))
            for j in range(k + 1):
                acc = acc + M[k, j] * X[:, j]
            Y[:, k] = acc
    elif mode == "upper":
        for k in range(n1):
            acc = np.zeros(len(X))
            for j in range(k, n1):
                acc = acc + M[k, j] * X[:, j]
            Y[:, k] = acc
    elif mode == "solve":
        for k in range(n1):
            s = np.zeros(len(X))
            for j in range(k):
                s = s + M[k, j] * Y[:, j]
            Y[:, k] = (X[:, k] - s) / M[k, k]
    return Y


def sweep(x, tabs, h, M, mode, n1):
    idx = thread_index_matrix(tabs, h, n1)
    valid = idx >= 0
    X = np.where(valid, x[np.maximum(idx, 0)], 0.0)
    Y = apply_tri(X, M, mode)
    out = np.empty_like(x)
    out[:] = np.nan
    out[idx[valid]] = Y[valid]
    assert not np.isnan(out).any(), "tables do not cover every element exactly once"
    return out


def whole_function(x, tabs: nat.IndexTables, M, n1, mode="lower"):
    """Emulates k_whole (lpfun_b200/csrc/whole_kernel.cuh) from its REAL item tables: the function is staged, the
    coordinates 0 .. m-2 run in place over the item lists (rounds of `threads` items in thread order), the last
    coordinate reads its fibres and writes the result."""
    meta = tabs.whole_meta()
    assert meta["ok"], "set is not eligible for the whole-function kernel"
    items = tabs.table(nat.TAB_WHOLE_ITEMS).astype(np.int64)
    seo = tabs.table(nat.TAB_WHOLE_SEO).astype(np.int64)
    m = tabs.m
    TG = meta["threads"]
    kk = np.arange(n1)[None, :]
    buf = x.copy()
    for h in range(m):
        lst = items[meta["item_off"][h]: meta["item_off"][h] + meta["n_items"][h]]
        assert len(lst) % TG == 0
        t = (lst >> 16) & 63
        lo = lst & 0xFFFF
        lane = lst >> 22
        live = t > 0
        t, lo, lane = t[live], lo[live], lane[live]
        valid = kk < t[:, None]
        if h == 0:
            idx = lo[:, None] + kk
        else:
            idx = seo[np.where(valid, lo[:, None] + kk, 0)] + lane[:, None]
        flat = idx[valid]
        assert len(np.unique(flat)) == len(flat) == len(x), "items do not cover every element exactly once"
        X = np.where(valid, buf[np.where(valid, idx, 0)], 0.0)
        Y = apply_tri(X, M, mode)
        buf[idx[valid]] = Y[valid]
    return buf


def whole_function_split(x, tabs: nat.IndexTables, M, n1):
    """Emulates k_whole_split (lower mat-vec, m = 3) from its REAL tables: unit P (planes a_2 < C) sweeps coordinates 0, 1
    in place, exports the coordinate-1 results of the columns that reach beyond C to the scratch [a_2][column], folds the rows
    k < C of the last coordinate; unit Q sweeps its planes, then folds the rows k >= C from the first term on (j < C from the
    scratch, j >= C from its buffer)."""
    meta = tabs.wsplit_meta()
    assert meta["ok"], "set is not eligible for the split whole-function kernel"
    items = tabs.table(nat.TAB_WSPLIT_ITEMS).astype(np.int64)
    seo = tabs.table(nat.TAB_WSPLIT_SEO).astype(np.int64)
    C, ncolq, cb, rq = meta["C"], meta["ncolq"], np.array(meta["cb"]), np.array(meta["rq"])
    out = np.full_like(x, np.nan)
    scratch = np.full((C, max(ncolq, 1)), np.nan)
    kk = np.arange(n1)[None, :]
    seen = np.zeros(len(x), dtype=int)
    for p in range(2):
        e0, ne = meta["e0"][p], meta["ne"][p]
        buf = x[e0:e0 + ne].copy()
        for h in range(3):
            lst = items[meta["item_off"][p][h]: meta["item_off"][p][h] + meta["n_items"][p][h]]
            t = (lst >> 16) & 63
            live = t > 0
            lst, t = lst[live], t[live]
            lo, lane, aux = lst & 0xFFFF, (lst >> 22) & 31, (lst >> 27) & 31
            if h == 0:
                valid = kk < t[:, None]
                idx = lo[:, None] + kk
                X = np.where(valid, buf[np.where(valid, idx, 0)], 0.0)
                Y = apply_tri(X, M, "lower")
                assert len(np.unique(idx[valid])) == ne
                buf[idx[valid]] = Y[valid]
            elif h == 1:
                valid = kk < t[:, None]
                idx = seo[np.where(valid, lo[:, None] + kk, 0)] + lane[:, None]
                X = np.where(valid, buf[np.where(valid, idx, 0)], 0.0)
                Y = apply_tri(X, M, "lower")
                assert len(np.unique(idx[valid])) == ne
                buf[idx[valid]] = Y[valid]
                if p == 0:  # export: element k of the fibre is (a_0 = lane, a_1 = k, a_2 = aux)
                    exp = valid & (lane[:, None] < rq[np.minimum(kk, 31)])
                    rows, cols = np.nonzero(exp)
                    scratch[aux[rows], cb[cols] + lane[rows]] = Y[rows, cols]
            elif p == 0:  # rows k < C of every column
                valid = kk < t[:, None]
                idx = seo[np.where(valid, lo[:, None] + kk, 0)] + lane[:, None]
                X = np.where(valid, buf[np.where(valid, idx, 0)], 0.0)
                Y = apply_tri(X, M, "lower")
                out[e0 + idx[valid]] = Y[valid]
                np.add.at(seen, e0 + idx[valid], 1)
            else:  # rows k >= C: x[j < C] from the scratch, x[j >= C] from the buffer (seo lists start at plane C)
                assert np.all(t > C)
                col = cb[aux] + lane
                assert np.all(lane < rq[aux])
                valid_hi = (kk >= C) & (kk < t[:, None])
                idx = seo[np.where(valid_hi, lo[:, None] + kk, 0)] + lane[:, None]  # the lists of part Q start with C placeholders
                X = np.where(valid_hi, buf[np.where(valid_hi, idx, 0)], 0.0)
                X[:, :C] = scratch[:, col].T
                assert not np.isnan(X[:, :C]).any(), "scratch entry missing"
                Y = apply_tri(X, M, "lower")
                out[e0 + idx[valid_hi]] = Y[valid_hi]
                np.add.at(seen, e0 + idx[valid_hi], 1)
    assert np.all(seen == 1), "outputs are not covered exactly once"
    return out
