"""Host-side plan build of the product (lpfun_b200.utils, lpfun_b200.core.set via the C++ host
code in liblpfnt.so) against the reference's tables/matrices in tests/golden/.  Bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_case


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_lp_set_tube_entropy_match_reference(built_lib):
    from lpfun_b200.core import set as S

    rows = json.load(open(os.path.join(GOLDEN, "tables.json")))
    for r in rows:
        p = np.inf if r["p"] == "inf" else r["p"]
        A = S.lp_set(r["m"], r["n"], p)
        assert A.dtype == np.int64 and A.shape == (r["N"], r["m"])
        assert sha(A) == r["A"], r
        T = S.lp_tube(A)
        assert sha(T) == r["T"], r
        assert list(S.entropy(A)) == r["e_T"], r


def test_tube_cardinalities(built_lib):
    """The reference's own table tests (test.py:19-42)."""
    from itertools import product

    from lpfun_b200.core import set as S
    from lpfun_b200.core.utils import binomial

    for m in range(1, 7):
        for n in (4, 5, 6, 7, 8):
            assert S.lp_tube(S.lp_set(m, n, 1.0)).sum() == binomial(n + m, m)
            if m <= 4:
                card = sum(1 for pt in product(range(n + 1), repeat=m) if np.linalg.norm(pt) <= n)
                assert S.lp_tube(S.lp_set(m, n, 2.0)).sum() == card


def test_c3_tables(built_lib):
    from lpfun_b200.core import set as S

    g = np.load(os.path.join(GOLDEN, "c3_d6n20.npz"))
    A = S.lp_set(6, 20, 2.0)
    assert len(A) == 6_974_412 and sha(A) == str(g["A_sha"])
    assert sha(S.lp_tube(A)) == str(g["T_sha"])
    assert list(S.entropy(A)) == list(g["e_T"])


@pytest.mark.parametrize("name", golden_cases("newton"))
def test_nodes_and_matrices_bit_exact(name):
    from lpfun_b200 import utils as U

    g = load_case(name)
    n = int(g["n"])
    x = (U.leja_nodes if str(g["nodes_kind"]) == "leja" else U.cheb2nd_nodes)(n + 1)
    lo = U.get_leja_order(x)
    assert np.array_equal(lo, g["leja_order"])
    xs = x[lo]
    assert np.array_equal(xs, g["x"])
    V, D = U.newton2lagrange(xs), U.newton2derivative(xs)
    D2 = D @ D
    D3 = D @ D2
    assert np.array_equal(V, g["Vx"]) and np.array_equal(D, g["Dx"])
    assert np.array_equal(D2, g["Dx2"]) and np.array_equal(D3, g["Dx3"])
    assert np.array_equal(U.pack_lower(V), g["Vx_lt"])
    if bool(g["precomputation"]):
        assert np.array_equal(U.pack_lower(np.linalg.inv(V)), g["inv_Vx_lt"])
    for q, M in enumerate((D, D2, D3)):
        assert np.array_equal(U.pack_upper(M), g[f"Dx_lt{q}"])
        assert np.array_equal(U.pack_lower(M.T), g[f"DxT_lt{q}"])
    assert U.is_lower_triangular(V) and U.is_lower_triangular(D.T)


@pytest.mark.parametrize("name", golden_cases("chebyshev"))
def test_chebyshev_matrices_bit_exact(name):
    from lpfun_b200 import utils as U

    g = load_case(name)
    xs = g["x"]
    assert np.array_equal(U.chebyshev2lagrange(xs), g["Vx"])
    assert np.array_equal(U.chebyshev2derivative(xs), g["Dx"])
    L, Uu = U.get_lu(g["Vx"])
    assert np.array_equal(U.pack_lower(L), g["Vx_lt"])
    assert np.array_equal(U.pack_upper(Uu), g["Vx_ut"])
    if g["inv_Vx_lt"].size:
        assert np.array_equal(U.pack_lower(np.linalg.inv(L)), g["inv_Vx_lt"])
        assert np.array_equal(U.pack_upper(np.linalg.inv(Uu)), g["inv_Vx_ut"])
    D = g["Dx"]
    for q, M in enumerate((D, D @ D, D @ (D @ D))):
        assert np.array_equal(U.pack_upper(M), g[f"Dx_lt{q}"])
        assert np.array_equal(U.pack_lower(M.T), g[f"DxT_lt{q}"])
    assert not U.is_lower_triangular(g["Vx"]) and U.is_lower_triangular(D.T)


def test_classify_errors():
    from lpfun_b200 import utils as U

    with pytest.raises(ValueError):
        U.classify(0, 3, 2.0)
    with pytest.raises(ValueError):
        U.classify(2, 3, 2.5)
    with pytest.raises(ValueError):
        U.classify(2, -1, 2.0)
    assert U.classify(2, 3, np.inf)
    assert np.array_equal(U.cheb2nd_nodes(1), [-1.0, 1.0]) and np.array_equal(U.leja_nodes(0), [0.0])


def test_ordinal_embedding(built_lib):
    from lpfun_b200.core import set as S

    small, big = S.lp_set(3, 4, 1.0), S.lp_set(3, 7, 2.0)
    phi = S.ordinal_embedding(small, big)
    assert np.array_equal(big[phi], small)
    with pytest.raises(ValueError):
        S.ordinal_embedding(big, small)
