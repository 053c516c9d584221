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
    """Emulates k_sweep_tiles: per tile the rows go to a padded buffer, an optional sweep along
    coordinate 0 over the rows (sorted row list), then the sweep along h over the sorted items."""
    tile_fib = tabs.table(nat.TAB_TILE_FIB, h)
    ptr = tabs.table(nat.TAB_FIB_PTR, h)
    ent = tabs.table(nat.TAB_FIB_ENT, h)
    tstart = tabs.table(nat.TAB_THR_START, h)
    items = tabs.table(nat.TAB_TILE_ITEMS, h)
    rows0 = tabs.table(nat.TAB_TILE_ROWS0, h)
    out = np.full_like(x, np.nan)
    PAD = n1 + 1
    for tl in range(len(tile_fib) - 1):
        f0, f1 = tile_fib[tl], tile_fib[tl + 1]
        r0, r1 = ptr[f0], ptr[f1]
        i0, i1 = tstart[f0], tstart[f1]
        R = r1 - r0
        assert R <= 128 and i1 - i0 <= 128
        buf = np.zeros((R, PAD))
        for r in range(R):
            off, ln = ent[r0 + r]
            buf[r, :ln] = x[off:off + ln]
        if dim0:
            d = rows0[r0:r1]
            rr, tt = (d & 0xFFFF).astype(int), (d >> 24).astype(int)
            assert sorted(rr) == list(range(R)) and np.all(np.diff(tt) <= 0)
            X = np.where(np.arange(n1)[None, :] < tt[:, None], buf[rr][:, :n1], 0.0)
            Y = apply_tri(X, M, mode)
            for q, (r, t) in enumerate(zip(rr, tt)):
                buf[r, :t] = Y[q, :t]
        d = items[i0:i1]
        rr, ll, tt = (d & 0xFFFF).astype(int), ((d >> 16) & 0xFF).astype(int), (d >> 24).astype(int)
        assert np.all(np.diff(tt) <= 0)
        X = np.zeros((len(d), n1))
        for q, (r, l, t) in enumerate(zip(rr, ll, tt)):
            X[q, :t] = buf[r:r + t, l]
        Y = apply_tri(X, M, mode)
        for q, (r, l, t) in enumerate(zip(rr, ll, tt)):
            buf[r:r + t, l] = Y[q, :t]
        for r in range(R):
            off, ln = ent[r0 + r]
            out[off:off + ln] = buf[r, :ln]
    assert not np.isnan(out).any()
    return out
