"""Validates the derived device tables (fibre CSR, thread maps) and the per-thread fold order by
emulating the GPU decomposition in NumPy and comparing with the reference outputs -- bit-exact.
No GPU needed."""
import numpy as np
import pytest

from conftest import golden_cases, load_case
from lpfun_b200 import _native as nat
from lpfun_b200 import utils as U
import sim

CASES = [c for c in golden_cases("newton", small_only=True) if c != "c2_d3n20_leja"] + ["c2_d3n20_leja"]


@pytest.mark.parametrize("name", CASES)
def test_emulated_gpu_decomposition_matches_reference(name, built_lib):
    g = load_case(name)
    m, n = int(g["m"]), int(g["n"])
    n1 = n + 1
    A = g["A"].astype(np.int64)
    tabs = nat.IndexTables(A, n)
    P = g["P"] if bool(g["colex"]) else None
    # every coordinate: threads cover each element exactly once, lines never grow along a fibre
    for h in range(m):
        idx = sim.thread_index_matrix(tabs, h, n1)
        flat = idx[idx >= 0]
        assert len(flat) == len(A) and len(np.unique(flat)) == len(A)
        # the k-th entry of a fibre has a_h == k and all other coordinates equal
        rows, cols = np.nonzero(idx >= 0)
        assert np.array_equal(A[idx[rows, cols], h], cols)
        first = idx[rows, 0]
        other = [j for j in range(m) if j != h]
        assert np.array_equal(A[idx[rows, cols]][:, other], A[first][:, other])
    # fnt (default: inv V mat-vec, or forward substitution), then ifnt
    v = g["values"][0]
    x = v[P] if P is not None else v.copy()
    if bool(g["precomputation"]):
        M, mode = U.unpack_lower(g["inv_Vx_lt"], n1), "lower"
    else:
        M, mode = U.unpack_lower(g["Vx_lt"], n1), "solve"
    for h in range(m):
        x = sim.sweep(x, tabs, h, M, mode, n1)
    assert np.array_equal(x, g["fnt"][0])
    V = U.unpack_lower(g["Vx_lt"], n1)
    y = g["fnt"][0].copy()
    for h in range(m):
        y = sim.sweep(y, tabs, h, V, "lower", n1)
    if P is not None:
        z = np.empty_like(y)
        z[P] = y
        y = z
    assert np.array_equal(y, g["ifnt"][0])
    # dx / dxT along every coordinate
    for i in range(m):
        key = f"dx_{i}_1"
        if key in g.files and not (np.isinf(float(g["p"])) and m > 1):
            D = U.unpack_upper(g["Dx_lt0"], n1)
            assert np.array_equal(sim.sweep(g["fnt"][0], tabs, i, D, "upper", n1), g[key])
        key = f"dxT_{i}_2"
        if key in g.files:
            DT = U.unpack_lower(g["DxT_lt1"], n1)
            assert np.array_equal(sim.sweep(g["rnd"], tabs, i, DT, "lower", n1), g[key])
    # eval line descriptors
    if m >= 2:
        al = tabs.table(nat.TAB_LINE_ALPHA).reshape(-1, m - 1)
        off = tabs.table(nat.TAB_LINE_OFF)
        assert np.array_equal(al, A[off[:-1], 1:])
    assert list(tabs.table(nat.TAB_LEVEL_SIZE)) == list(g["e_T"]) or len(A) == 1


def test_index_rejects_bad_sets(built_lib):
    A = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [2, 1]], dtype=np.int64)  # not downward closed
    with pytest.raises(ValueError):
        nat.IndexTables(A, 2)
    B = np.array([[0, 0], [2, 0]], dtype=np.int64)  # coordinate 0 skips a value
    with pytest.raises(ValueError):
        nat.IndexTables(B, 2)


@pytest.mark.parametrize("threads,fibres", [(384, 1), (768, 1), (384, 2)])
@pytest.mark.parametrize("name", ["c1_d2n10", "d3n12_p1", "d3n6_inf", "d3n12_solve", "c2_d3n20_leja", "c4_d3n30", "d2n10_solve"])
def test_emulated_whole_function_kernel_matches_reference(name, threads, fibres, built_lib):
    """k_whole from its real item tables (rounds of 384 or 768 items in thread order, one or two per thread) -- bit-exact."""
    g = load_case(name)
    m, n = int(g["m"]), int(g["n"])
    n1 = n + 1
    tabs = nat.IndexTables(g["A"].astype(np.int64), n)
    tabs.whole_build(threads=threads, fibres=fibres)
    meta = tabs.whole_meta()
    assert meta["ok"] and meta["threads"] == threads
    if name == "c4_d3n30":
        assert meta["max_items"] == 736 and meta["n_items"] == ([768] * 3)
    P = g["P"] if bool(g["colex"]) else None
    v = g["values"][0]
    x = v[P] if P is not None else v.copy()
    if bool(g["precomputation"]):
        M, mode = U.unpack_lower(g["inv_Vx_lt"], n1), "lower"
    else:
        M, mode = U.unpack_lower(g["Vx_lt"], n1), "solve"
    assert np.array_equal(sim.whole_function(x, tabs, M, n1, mode), g["fnt"][0])
    V = U.unpack_lower(g["Vx_lt"], n1)
    y = sim.whole_function(g["fnt"][0], tabs, V, n1, "lower")
    if P is not None:
        z = np.empty_like(y)
        z[P] = y
        y = z
    assert np.array_equal(y, g["ifnt"][0])


@pytest.mark.parametrize("C", [0, 5, 13])
@pytest.mark.parametrize("name", ["c4_d3n30", "c2_d3n20_leja", "d3n12_p1"])
def test_emulated_split_whole_function_kernel_matches_reference(name, C, built_lib):
    """k_whole_split (two units per function, so that two CTAs fit on an SM) from its real tables -- bit-exact, including the
    hand-over of the coordinate-1 results of part P to the last sweep of part Q."""
    g = load_case(name)
    n = int(g["n"])
    n1 = n + 1
    tabs = nat.IndexTables(g["A"].astype(np.int64), n)
    tabs.wsplit_build(C=min(C, n))
    meta = tabs.wsplit_meta()
    assert meta["ok"] and sum(meta["ne"]) == len(g["A"])
    P = g["P"] if bool(g["colex"]) else None
    v = g["values"][0]
    x = v[P] if P is not None else v.copy()
    assert bool(g["precomputation"])
    M = U.unpack_lower(g["inv_Vx_lt"], n1)
    assert np.array_equal(sim.whole_function_split(x, tabs, M, n1), g["fnt"][0])
    V = U.unpack_lower(g["Vx_lt"], n1)
    y = sim.whole_function_split(g["fnt"][0], tabs, V, n1)
    if P is not None:
        z = np.empty_like(y)
        z[P] = y
        y = z
    assert np.array_equal(y, g["ifnt"][0])
