"""Small helpers mirroring src/lpfun/core/utils.py."""
from __future__ import annotations

import math

import numpy as np


def apply_permutation(P: np.ndarray, x: np.ndarray, invert: bool = False) -> np.ndarray:
    """``invert=True`` gathers (y[i] = x[P[i]]), otherwise scatters (y[P[i]] = x[i]).

    Reference: src/lpfun/core/utils.py:6-20.  Works on vectors and on row-stacked arrays
    (the grid), like the reference.
    """
    P = np.asarray(P, dtype=np.int64)
    x = np.asarray(x)
    if invert:
        return x[P].copy()
    y = np.zeros_like(x)
    y[P] = x
    return y


def binomial(n: int, m: int) -> int:
    """Exact binomial coefficient (src/lpfun/core/utils.py:23-30)."""
    n, m = int(n), int(m)
    if m < 0 or m > n:
        return 0
    return math.comb(n, m)


def memory_estimate(m: int, n: int, p: float) -> int:
    """Upper bound on |A| the reference uses to pre-allocate (src/lpfun/core/utils.py:33-46).

    Kept for API parity; the native enumerator counts first and needs no estimate."""
    m, n, p = int(m), int(n), float(p)
    if p <= 1.0:
        singular = 1 + m * n
        b = binomial(n + m, m) if m < n else binomial(n + m, n)
        return int(singular * (1 - p) + b * p)
    if p <= 2.0:
        fac1 = (p * math.e / m) ** (1 / p)
        fac2 = math.sqrt(p / (2 * math.pi * m))
        return int(math.ceil((fac1 * (n + 2) * math.gamma(1 + 1 / p)) ** m * fac2))
    return int((n + 1) ** m)
