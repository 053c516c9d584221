#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Usage (in the build container, where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py [--only NAME] [--skip-c3]

The reference (phil-hofmann/lpFun, pure Python + Numba) is imported from
/root/reference/src (or $LPFUN_REF).  Nothing from the reference is copied into
the repo: only its *outputs* on seeded inputs are stored, as compressed .npz
files, together with this script.  The GPU box has no /root/reference; tests
read only the fixtures.

Fixtures
--------
tables.json         sha256 + sizes of A / T / e_T for the whole test.py grid
                    (m=1..6, p in {1,2,inf}, n=4..8) plus fractional p.
case_<name>.npz     one Transform each: tables (A,T,P,leja order, nodes, packed
                    matrices) and fnt / ifnt / dx / dxT / eval outputs.
c3_d6n20.npz        config 3 (N = 6 974 412): table hashes + sampled outputs.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LPFUN_REF", "/root/reference/src")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import lpfun  # noqa: E402  (the reference)
from lpfun import Transform  # noqa: E402
from lpfun.utils import leja_nodes, cheb2nd_nodes  # noqa: E402

SEED = 1234


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------
# smooth test functions of SURVEY.md §8(d)
# ---------------------------------------------------------------------------


def f_c1(g):
    return np.sin(g[:, 0]) * np.cos(g[:, 1])


def f_c2(g):
    return np.sin(g[:, 0]) + np.cos(g[:, 1]) + np.exp(g[:, 2])


def f_smooth(g):
    m = g.shape[1]
    w = np.arange(1, m + 1, dtype=np.float64) / 6.0
    return np.exp(-np.sum(g * g, axis=1)) * np.sin(g @ w)


def f_sines(g, rng, q=4):
    m = g.shape[1]
    a = rng.uniform(0.5, 1.0, size=q)
    w = rng.uniform(-3.0, 3.0, size=(q, m))
    ph = rng.uniform(0.0, 2 * np.pi, size=q)
    return np.sum(a[None, :] * np.sin(g @ w.T + ph[None, :]), axis=1)


FUNCS = {"c1": f_c1, "c2": f_c2, "smooth": f_smooth}

# name, m, n, p, nodes, basis, precomputation, colex_order, function, n_eval_points
CASES = [
    ("c1_d2n10", 2, 10, 2.0, "cheb2nd", "newton", True, True, "c1", 10),
    ("c2_d3n20_leja", 3, 20, 2.0, "leja", "newton", True, True, "c2", 10),
    ("c4_d3n30", 3, 30, 2.0, "cheb2nd", "newton", True, True, "sines", 16),
    ("c5_d4n20", 4, 20, 2.0, "cheb2nd", "newton", True, True, "smooth", 64),
    ("d1n12", 1, 12, 2.0, "cheb2nd", "newton", True, True, "smooth", 8),
    ("d2n7_inf", 2, 7, np.inf, "cheb2nd", "newton", True, True, "smooth", 8),
    ("d3n6_inf", 3, 6, np.inf, "leja", "newton", True, False, "smooth", 8),
    ("d3n12_p1", 3, 12, 1.0, "leja", "newton", True, True, "smooth", 8),
    ("d4n8", 4, 8, 2.0, "cheb2nd", "newton", True, False, "smooth", 8),
    ("d5n5_p1", 5, 5, 1.0, "cheb2nd", "newton", True, True, "smooth", 8),
    ("d5n6", 5, 6, 2.0, "leja", "newton", True, True, "smooth", 8),
    ("d6n4", 6, 4, 2.0, "cheb2nd", "newton", True, True, "smooth", 8),
    ("d3n9_p05", 3, 9, 0.5, "cheb2nd", "newton", True, True, "smooth", 8),
    ("d4n7_p15", 4, 7, 1.5, "leja", "newton", True, True, "smooth", 8),
    # precomputation=False: forward-substitution path (transform_lt_*)
    ("d2n10_solve", 2, 10, 2.0, "cheb2nd", "newton", False, True, "c1", 8),
    ("d3n12_solve", 3, 12, 2.0, "leja", "newton", False, True, "c2", 8),
    ("d4n8_solve", 4, 8, 2.0, "cheb2nd", "newton", False, False, "smooth", 8),
    ("d1n9_solve", 1, 9, 2.0, "cheb2nd", "newton", False, True, "smooth", 8),
    ("d3n5_inf_solve", 3, 5, np.inf, "cheb2nd", "newton", False, True, "smooth", 8),
    # Chebyshev basis (LU two-pass), both precomputation modes
    ("d2n8_cheb", 2, 8, 2.0, "cheb2nd", "chebyshev", True, True, "c1", 8),
    ("d3n7_cheb", 3, 7, 2.0, "cheb2nd", "chebyshev", True, False, "c2", 8),
    ("d4n6_cheb", 4, 6, 1.0, "cheb2nd", "chebyshev", True, True, "smooth", 8),
    ("d3n7_cheb_solve", 3, 7, 2.0, "cheb2nd", "chebyshev", False, True, "c2", 8),
    ("d1n8_cheb", 1, 8, 2.0, "cheb2nd", "chebyshev", True, True, "smooth", 8),
    ("d2n6_inf_cheb", 2, 6, np.inf, "cheb2nd", "chebyshev", True, True, "smooth", 8),
    ("d4n5_cheb_solve", 4, 5, 2.0, "cheb2nd", "chebyshev", False, False, "smooth", 8),
]


def build(m, n, p, nodes, basis, pre, colex):
    return Transform(
        m,
        n,
        p,
        nodes=leja_nodes if nodes == "leja" else cheb2nd_nodes,
        basis=basis,
        precomputation=pre,
        precompilation=False,
        colex_order=colex,
        report=False,
    )


def opt(a):
    return np.zeros(0) if a is None else np.asarray(a)


def dump_case(name, m, n, p, nodes, basis, pre, colex, fname, npts):
    t0 = time.time()
    t = build(m, n, p, nodes, basis, pre, colex)
    N = len(t)
    rng = np.random.default_rng(SEED)
    if fname == "sines":
        vals = np.stack([f_sines(t.grid, rng), f_sines(t.grid, rng)])
    else:
        vals = FUNCS[fname](t.grid)[None, :]
    rnd = rng.standard_normal(N)
    pts = rng.random((npts, m)) * 2 - 1
    out = dict(
        m=m, n=n, p=p, nodes_kind=nodes, basis=basis, precomputation=pre, colex=colex,
        N=N,
        A=t.multi_index_set.astype(np.uint8 if n < 256 else np.int64),
        T=t.tube,
        e_T=np.asarray(lpfun.core.set.entropy(t.tube)),
        P=opt(t.colex_order),
        leja_order=t.leja_order,
        x=t.nodes,
        grid_sha=sha(t.grid),
        Vx=t._Vx, Dx=t._Dx, Dx2=t._Dx2, Dx3=t._Dx3,
        cond_Vx=t._cond_Vx,
        Vx_lt=opt(t._Vx_lt), Vx_ut=opt(t._Vx_ut),
        inv_Vx_lt=opt(t._inv_Vx_lt), inv_Vx_ut=opt(t._inv_Vx_ut),
        values=vals, rnd=rnd, points=pts,
    )
    for q in range(3):
        out[f"Dx_lt{q}"] = opt(t._Dx_lt[q])
        out[f"DxT_lt{q}"] = opt(t._DxT_lt[q])
        out[f"Dx_ut{q}"] = opt(t._Dx_ut[q]) if t._Dx_ut is not None else np.zeros(0)
        out[f"DxT_ut{q}"] = opt(t._DxT_ut[q]) if t._DxT_ut is not None else np.zeros(0)
    coeffs = np.stack([t.fnt(v) for v in vals])
    out["fnt"] = coeffs
    out["ifnt"] = np.stack([t.ifnt(c) for c in coeffs])
    out["fnt_rnd"] = t.fnt(rnd)
    out["ifnt_rnd"] = t.ifnt(rnd)
    c0 = coeffs[0]
    for i in range(m):
        for k in (1, 2, 3):
            # store all (i,k) for small sets, only k=1 (+ i=0 all k) for the big ones
            if N > 20000 and not (k == 1 or i == 0):
                continue
            out[f"dx_{i}_{k}"] = t.dx(c0, i, k)
            out[f"dxT_{i}_{k}"] = t.dxT(rnd, i, k)
    out["eval"] = t.eval(c0, pts)
    out["eval_dx0"] = t.eval(out["dx_0_1"], pts)
    np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
    print(f"  {name}: N={N} ({time.time()-t0:.1f}s)", flush=True)


def dump_tables():
    rows = []
    grid = [(m, n, p) for m in range(1, 7) for p in (1.0, 2.0, np.inf) for n in (4, 5, 6, 7, 8)]
    grid += [(m, n, p) for m in (2, 3, 4) for p in (0.3, 0.5, 0.7, 1.3, 1.5, 1.9) for n in (3, 6, 9, 11)]
    grid += [(2, 10, 2.0), (3, 20, 2.0), (3, 30, 2.0), (4, 20, 2.0), (5, 20, 2.0), (2, 100, 2.0), (3, 40, 1.0)]
    for m, n, p in grid:
        if p == np.inf and (n + 1) ** m > 600000:
            continue
        A = lpfun.core.set.lp_set(m, n, p)
        T = lpfun.core.set.lp_tube(A, m, n, p)
        e_T = lpfun.core.set.entropy(T)
        rows.append(dict(m=m, n=n, p=("inf" if p == np.inf else p), N=int(len(A)), N1=int(len(T)),
                         A=sha(A.astype(np.int64)), T=sha(T.astype(np.int64)),
                         e_T=[int(v) for v in e_T]))
    with open(os.path.join(HERE, "tables.json"), "w") as fh:
        json.dump(rows, fh, indent=0)
    print(f"  tables.json: {len(rows)} rows", flush=True)


def dump_c3():
    """Config 3: Transform(6, 20).  Too large to store: hashes + sampled outputs."""
    t0 = time.time()
    t = Transform(6, 20, report=False, precompilation=False)
    N = len(t)
    vals = f_smooth(t.grid)
    c = t.fnt(vals)
    r = t.ifnt(c)
    rng = np.random.default_rng(SEED)
    idx = np.sort(rng.choice(N, size=16384, replace=False))
    out = dict(
        N=N, A_sha=sha(t.multi_index_set), T_sha=sha(t.tube), P_sha=sha(t.colex_order),
        grid_sha=sha(t.grid), e_T=np.asarray(lpfun.core.set.entropy(t.tube)),
        idx=idx, values_s=vals[idx], fnt_s=c[idx], ifnt_s=r[idx],
        fnt_absmax=np.max(np.abs(c)), ifnt_absmax=np.max(np.abs(r)),
        fnt_sha=sha(c), ifnt_sha=sha(r), values_sha=sha(vals),
        fnt_sum=np.sum(c), roundtrip_err=np.max(np.abs(r - vals)),
    )
    np.savez_compressed(os.path.join(HERE, "c3_d6n20.npz"), **out)
    print(f"  c3_d6n20: N={N} ({time.time()-t0:.1f}s)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--skip-tables", action="store_true")
    args = ap.parse_args()
    import numba

    print(f"reference at {REF}; numba {numba.__version__} threads {numba.get_num_threads()}; numpy {np.__version__}")
    if not args.skip_tables and args.only is None:
        dump_tables()
    for case in CASES:
        if args.only and case[0] != args.only:
            continue
        dump_case(*case)
    if not args.skip_c3 and args.only in (None, "c3"):
        dump_c3()
