"""Host-side 1-D numerics of the plan build (NumPy, float64).

These are the small O(n^2)..O(n^3) pieces the reference computes once per
``Transform`` (reference: src/lpfun/utils.py).  They run on the host because
the packed matrices must match the reference bit-for-bit (SURVEY.md §7 "hard
parts"): every function below performs the reference's floating-point
operations in the reference's order, and ``np.linalg.inv`` / ``@`` are the very
same NumPy calls the reference makes (src/lpfun/transform.py:276-277,288).

Nothing here touches the GPU; the device kernels only ever see the packed
outputs of :func:`pack_lower` / :func:`pack_upper_as_lower_T`.
"""

from __future__ import annotations

import math

import numpy as np

__all__ = [
    "classify",
    "cheb2nd_nodes",
    "leja_nodes",
    "get_leja_order",
    "newton2lagrange",
    "newton2derivative",
    "chebyshev2lagrange",
    "chebyshev2derivative",
    "is_lower_triangular",
    "get_rmo",
    "get_lu",
    "pack_lower",
    "pack_upper",
    "unpack_lower",
    "unpack_upper",
]


def classify(m: int, n: int, p: float) -> bool:
    """Argument validation, same messages/conditions as src/lpfun/utils.py:10-21."""
    m, n, p = int(m), int(n), float(p)
    if m < 1:
        raise ValueError("The parameter dim should be at least 1.")
    if (p <= 0.0 or p > 2.0) and not p == np.inf:
        raise ValueError("The parameter p should be in the range (0, 2] or inf.")
    if n < 0:
        raise ValueError("The parameter degree should be non-negative.")
    return True


# ----------------------------------------------------------------------------
# nodes
# ----------------------------------------------------------------------------


def cheb2nd_nodes(n: int) -> np.ndarray:
    """``n`` Chebyshev points of the 2nd kind, cos(k*pi/(n-1)), k=0..n-1.

    Reference: src/lpfun/utils.py:27-37 (same NumPy expression so that the
    nodes are bit-identical; the special cases n=0,1 are the reference's).
    """
    n = int(n)
    if n < 0:
        raise ValueError("The parameter ``n`` should be non-negative.")
    if n == 0:
        return np.zeros(1, dtype=np.float64)
    if n == 1:
        return np.array([-1.0, 1.0], dtype=np.float64)
    k = np.arange(n, dtype=np.float64)
    return np.cos(k * np.pi / (n - 1))


def get_leja_order(nodes: np.ndarray, limit: int = -1) -> np.ndarray:
    """Greedy Leja ordering (reference: src/lpfun/utils.py:201-225).

    Start at index 0; each step picks, among the not-yet-chosen indices scanned
    in increasing order, the LAST index maximising ``|prod_j (x[lj[j]] - x[i])|``
    (comparison is ``>=`` against a running maximum that starts at 0), where
    the product is accumulated left to right starting from 1.
    """
    x = np.asarray(nodes, dtype=np.float64)
    n = len(x)
    limit = n if limit == -1 else int(limit)
    remaining = list(range(1, n))
    chosen = [0]
    xs = x.tolist()  # python floats: IEEE double arithmetic, same rounding
    for _ in range(limit - 1):
        best_pos, best = 0, 0.0
        for pos, i in enumerate(remaining):
            prod = 1.0
            xi = xs[i]
            for j in chosen:
                prod = prod * (xs[j] - xi)
            prod = abs(prod)
            if prod >= best:
                best_pos, best = pos, prod
        chosen.append(remaining.pop(best_pos))
    return np.asarray(chosen[:limit], dtype=np.int64)


def leja_nodes(n: int, m: int = 1000) -> np.ndarray:
    """First ``n`` Leja-ordered points of an ``m``-point cheb2nd sample.

    Reference: src/lpfun/utils.py:40-54.
    """
    n = int(n)
    if n < 0:
        raise ValueError("The parameter ``n`` should be non-negative.")
    if n == 0:
        return np.zeros(1, dtype=np.float64)
    if n == 1:
        return np.array([-1.0, 1.0], dtype=np.float64)
    if n > m:
        raise ValueError(
            f"The amount of nodes {n} must be smaller or equal than the sample size {m}."
        )
    sample = cheb2nd_nodes(m)
    return sample[get_leja_order(sample, limit=n)]


# ----------------------------------------------------------------------------
# 1-D matrices
# ----------------------------------------------------------------------------


def newton2lagrange(x: np.ndarray) -> np.ndarray:
    """Lower-triangular Newton Vandermonde V[i,j] = prod_{k<j} (x_i - x_k).

    Built column by column with the recurrence of src/lpfun/utils.py:60-73:
    V[i,j] = 1.0 * (V[i,j-1] * (x_i - x_{j-1})).
    """
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    V = np.zeros((n, n), dtype=np.float64)
    V[:, 0] = 1.0
    for j in range(1, n):
        V[:, j] = V[:, j - 1] * (x - x[j - 1])
    return V


def newton2derivative(x: np.ndarray) -> np.ndarray:
    """Strictly upper-triangular differentiation matrix in the Newton basis.

    Reference recurrence (src/lpfun/utils.py:95-110), on the transpose E = D^T:
    E[i,i-1] = i;  E[i,j] = (x_j - x_{i-1}) * E[i-1,j] + E[i-1,j-1]  for j<i-1,
    where E[i-1,-1] (j = 0) wraps to the last column, which is always 0.
    Multiply and add are rounded separately, as in the reference.
    """
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    E = np.zeros((n, n), dtype=np.float64)
    for i in range(1, n):
        E[i, i - 1] = float(i)
        if i >= 2:
            j = np.arange(0, i - 1)
            left = np.where(j >= 1, E[i - 1, j - 1], 0.0)
            E[i, : i - 1] = (x[: i - 1] - x[i - 1]) * E[i - 1, : i - 1] + left
    return np.ascontiguousarray(E.T)


def chebyshev2lagrange(x: np.ndarray) -> np.ndarray:
    """Chebyshev Vandermonde V[i,j] = T_j(x_i) by the three-term recurrence.

    Reference: src/lpfun/utils.py:76-89 (whose loop writes one column past the
    array; only the in-range columns are meaningful and are what we build).
    """
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    V = np.zeros((n, n), dtype=np.float64)
    V[:, 0] = 1.0
    if n > 1:
        V[:, 1] = x
    for j in range(2, n):
        V[:, j] = 2 * x * V[:, j - 1] - V[:, j - 2]
    return V


def chebyshev2derivative(x: np.ndarray) -> np.ndarray:
    """Differentiation matrix in the Chebyshev basis (node independent).

    Reference: src/lpfun/utils.py:113-126.
    """
    n = len(np.asarray(x)) - 1
    D = np.zeros((n + 1, n + 1), dtype=np.float64)
    for k in range(1, n + 1):
        D[k - 1 :: -2, k] = 2 * k
        D[0, k] *= 0.5
    return D


def is_lower_triangular(M: np.ndarray, atol: float = 1e-8) -> bool:
    """src/lpfun/utils.py:278-292: every strictly-upper entry has |.| < atol."""
    M = np.asarray(M, dtype=np.float64)
    iu = np.triu_indices(len(M), k=1)
    return bool(np.all(np.abs(M[iu]) < atol))


def get_lu(M: np.ndarray):
    """Doolittle LU without pivoting (src/lpfun/utils.py:315-328).

    Row operations are applied in the reference's order (column j outermost,
    multiplier L[i,j] = U[i,j]/U[j,j], then ``U[i,j:] -= L[i,j]*U[j,j:]``).
    The input is not modified (the reference's ``astype`` copies).
    """
    U = np.array(M, dtype=np.float64, copy=True)
    n = len(U)
    L = np.eye(n, dtype=np.float64)
    for j in range(n):
        for i in range(j + 1, n):
            L[i, j] = U[i, j] / U[j, j]
            U[i, j:] -= L[i, j] * U[j, j:]
    return L, U


# ----------------------------------------------------------------------------
# packing (row-major triangles)
# ----------------------------------------------------------------------------


def pack_lower(L: np.ndarray) -> np.ndarray:
    """Row-major packed lower triangle: row k at k(k+1)/2, k+1 entries.

    Same layout as the reference's ``get_rmo`` (src/lpfun/utils.py:295-309).
    """
    L = np.asarray(L, dtype=np.float64)
    return np.ascontiguousarray(L[np.tril_indices(len(L))])


get_rmo = pack_lower


def pack_upper(U: np.ndarray) -> np.ndarray:
    """Row-major packed upper triangle: row k at k*n1 - k(k-1)/2, n1-k entries.

    Equals the reference's double-flip idiom ``get_rmo(U[::-1, ::-1])[::-1]``
    (src/lpfun/transform.py:326).
    """
    U = np.asarray(U, dtype=np.float64)
    return np.ascontiguousarray(U[np.triu_indices(len(U))])


def unpack_lower(packed: np.ndarray, n1: int) -> np.ndarray:
    L = np.zeros((n1, n1), dtype=np.float64)
    L[np.tril_indices(n1)] = packed
    return L


def unpack_upper(packed: np.ndarray, n1: int) -> np.ndarray:
    U = np.zeros((n1, n1), dtype=np.float64)
    U[np.triu_indices(n1)] = packed
    return U


def pow_table(n: int, p: float) -> np.ndarray:
    """k**p for k=0..n through libm ``pow`` (the call Numba lowers ``**`` to in
    src/lpfun/core/set.py:23), one scalar at a time so no SIMD variant is used."""
    return np.array([math.pow(float(k), p) for k in range(n + 1)], dtype=np.float64)
