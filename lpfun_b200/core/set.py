"""l^p lower sets and their line/level structure (reference: src/lpfun/core/set.py).

The enumeration and the derived tables are built by the C++ host code of liblpfnt.so
(lpfun_b200/csrc/index_set.cpp); this module is the NumPy-facing wrapper with the reference's
function names and array conventions (int64, colexicographic order, coordinate 0 fastest).
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat


def lp_set(m: int, n: int, p: float) -> np.ndarray:
    """Multi-index set A = {a in N^m : ||a||_p <= n}, shape (N, m) int64, colex order.

    Reference: src/lpfun/core/set.py:58-74."""
    m, n, p = int(m), int(n), float(p)
    L = nat.lib()
    N = L.lpfnt_lp_set_count(m, n, p)
    if N < 0:
        nat.check(int(N))
    A = np.empty((N, m), dtype=np.int64)
    nat.check(L.lpfnt_lp_set_fill(m, n, p, A.ctypes.data, N))
    return A


def lp_tube(A: np.ndarray, m: int | None = None, n: int | None = None, p: float | None = None) -> np.ndarray:
    """Line lengths T along coordinate 0 (src/lpfun/core/set.py:96-107)."""
    A = np.ascontiguousarray(A, dtype=np.int64)
    if A.ndim != 2:
        raise ValueError("A must be (N, m)")
    L = nat.lib()
    n1 = L.lpfnt_lp_tube(A.ctypes.data, A.shape[0], A.shape[1], None)
    if n1 < 0:
        nat.check(int(n1))
    T = np.empty(n1, dtype=np.int64)
    L.lpfnt_lp_tube(A.ctypes.data, A.shape[0], A.shape[1], T.ctypes.data)
    return T


def entropy(A: np.ndarray) -> np.ndarray:
    """Level sizes [N_0, N_1, ..., n+1, 1] of the index trie.

    The reference derives them from T by self-indexing (src/lpfun/core/set.py:167-182), which
    relies on the coordinate symmetry of the set; here they are counted from A directly."""
    A = np.ascontiguousarray(A, dtype=np.int64)
    out = np.zeros(A.shape[1] + 1, dtype=np.int64)
    k = nat.lib().lpfnt_entropy(A.ctypes.data, A.shape[0], A.shape[1], out.ctypes.data)
    if k < 0:
        nat.check(k)
    return out[:k]


def ordinal_embedding(A_small: np.ndarray, A_big: np.ndarray) -> np.ndarray:
    """Positions of the members of ``A_small`` inside ``A_big`` (both colex ordered).

    Same result as src/lpfun/core/set.py:208-239 (which works on the tubes)."""
    a = np.ascontiguousarray(A_small, dtype=np.int64)
    b = np.ascontiguousarray(A_big, dtype=np.int64)
    if a.shape[1] != b.shape[1]:
        raise ValueError("Spatial dimensions do not match.")
    phi = np.empty(len(a), dtype=np.int64)
    nat.check(nat.lib().lpfnt_ordinal_embedding(a.ctypes.data, len(a), b.ctypes.data, len(b), a.shape[1], phi.ctypes.data))
    return phi
