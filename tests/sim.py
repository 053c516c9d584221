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


def sweep_tiles(x, tabs, h, M, mode, n1, dim0):
    """Emulates k_sweep_tiles: per tile the rows are staged PACKED (back to back), an optional sweep
    along coordinate 0 over the rows (sorted row list), then the sweep along h over the sorted items;
    copies in and out go through the element->row map exactly like the kernel."""
    tile_fib = tabs.table(nat.TAB_TILE_FIB, h)
    ptr = tabs.table(nat.TAB_FIB_PTR, h)
    ent = tabs.table(nat.TAB_FIB_ENT, h)
    tstart = tabs.table(nat.TAB_THR_START, h)
    items = tabs.table(nat.TAB_TILE_ITEMS, h)
    rows0 = tabs.table(nat.TAB_TILE_ROWS0, h)
    row_pos = tabs.table(nat.TAB_TILE_ROWPOS, h).astype(np.int64)
    elem_row = tabs.table(nat.TAB_TILE_ELEMROW, h).astype(np.int64)
    tile_elem = tabs.table(nat.TAB_TILE_ELEM, h)
    tile_nelem = tabs.table(nat.TAB_TILE_NELEM, h)
    out = np.full_like(x, np.nan)
    for tl in range(len(tile_fib) - 1):
        f0, f1 = tile_fib[tl], tile_fib[tl + 1]
        r0, r1 = ptr[f0], ptr[f1]
        i0, i1 = tstart[f0], tstart[f1]
        e0 = tile_elem[tl]
        e1 = e0 + tile_nelem[tl]
        assert e0 % 4 == 0
        R, E = r1 - r0, e1 - e0
        assert R <= 256 and i1 - i0 <= 256 and E <= 5120
        pos = row_pos[r0:r1]
        delta = ent[r0:r1, 0] - pos
        src = np.arange(E) + delta[elem_row[e0:e1]]
        buf = x[src].copy()
        if dim0:
            d = rows0[r0:r1]
            rr, tt = (d & 0xFFFF).astype(int), (d >> 24).astype(int)
            assert sorted(rr) == list(range(R)) and np.all(np.diff(tt) <= 0)
            idx = pos[rr][:, None] + np.arange(n1)[None, :]
            valid = np.arange(n1)[None, :] < tt[:, None]
            X = np.where(valid, buf[np.where(valid, idx, 0)], 0.0)
            Y = apply_tri(X, M, mode)
            buf[idx[valid]] = Y[valid]
        d = items[i0:i1]
        rr, ll, tt = (d & 0xFFFF).astype(int), ((d >> 16) & 0xFF).astype(int), (d >> 24).astype(int)
        assert np.all(np.diff(tt) <= 0)
        kk = np.arange(n1)[None, :]
        valid = kk < tt[:, None]
        rowidx = np.where(valid, rr[:, None] + kk, 0)
        idx = pos[rowidx] + ll[:, None]
        X = np.where(valid, buf[np.where(valid, idx, 0)], 0.0)
        Y = apply_tri(X, M, mode)
        buf[idx[valid]] = Y[valid]
        out[src] = buf
    assert not np.isnan(out).any()
    return out
