"""Pins the CPU oracle (oracle/lpfnt_oracle.c) against outputs of the UNMODIFIED reference
recorded in tests/golden/ (see tests/golden/gen_golden.py).  Everything here is bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_case
from oracle import oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_tables_match_reference_grid():
    rows = json.load(open(os.path.join(GOLDEN, "tables.json")))
    assert len(rows) >= 150
    for r in rows:
        p = np.inf if r["p"] == "inf" else r["p"]
        A = O.lp_set(r["m"], r["n"], p)
        T = O.lp_tube(A)
        assert len(A) == r["N"] and len(T) == r["N1"], r
        assert sha(A) == r["A"] and sha(T) == r["T"], r


def _oracle_for(g):
    mats = dict(
        Vx_lt=g["Vx_lt"],
        inv_Vx_lt=g["inv_Vx_lt"] if g["inv_Vx_lt"].size else None,
        Dx_ut=[g[f"Dx_lt{q}"] for q in range(3)],
        DxT_lt=[g[f"DxT_lt{q}"] for q in range(3)],
    )
    if g["Vx_ut"].size:  # LU-factorised V (Chebyshev basis)
        mats["Vx_ut"] = g["Vx_ut"]
        mats["inv_Vx_ut"] = g["inv_Vx_ut"] if g["inv_Vx_ut"].size else None
    return O.OracleTransform(int(g["m"]), int(g["n"]), float(g["p"]), nodes=str(g["nodes_kind"]),
                             precomputation=bool(g["precomputation"]), colex_order=bool(g["colex"]), mats=mats,
                             basis=str(g["basis"]))


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_bit_exact_vs_reference(name):
    g = load_case(name)
    t = _oracle_for(g)
    assert np.array_equal(t.A, g["A"].astype(np.int64))
    assert np.array_equal(t.T, g["T"])
    if bool(g["colex"]):
        assert np.array_equal(t.P, g["P"])
    assert sha(t.grid) == str(g["grid_sha"])
    assert np.array_equal(np.stack([t.fnt(v) for v in g["values"]]), g["fnt"])
    assert np.array_equal(np.stack([t.ifnt(c) for c in g["fnt"]]), g["ifnt"])
    assert np.array_equal(t.fnt(g["rnd"]), g["fnt_rnd"])
    assert np.array_equal(t.ifnt(g["rnd"]), g["ifnt_rnd"])
    for key in g.files:
        if key.startswith("dx_"):
            _, i, k = key.split("_")
            assert np.array_equal(t.dx(g["fnt"][0], int(i), int(k)), g[key]), key
        elif key.startswith("dxT_"):
            _, i, k = key.split("_")
            assert np.array_equal(t.dxT(g["rnd"], int(i), int(k)), g[key]), key
    assert np.array_equal(t.eval(g["fnt"][0], g["points"]), g["eval"])
    assert np.array_equal(t.eval(g["dx_0_1"], g["points"]), g["eval_dx0"])


def test_oracle_c3_sampled():
    """Config 3 (d=6, n=20, N=6 974 412): table hashes and sampled fnt/ifnt outputs."""
    g = np.load(os.path.join(GOLDEN, "c3_d6n20.npz"))
    A = O.lp_set(6, 20, 2.0)
    assert len(A) == int(g["N"]) and sha(A) == str(g["A_sha"])
    assert sha(O.lp_tube(A)) == str(g["T_sha"])
