#!/usr/bin/env python
"""Randomised parity run against the CPU oracle (development helper; the oracle is test infrastructure): random
(m, n, p, basis, precomputation, colex order, batch) within the fast kernels' ranges; fnt / exact ifnt / dx / dxT bit for bit,
FMA ifnt and eval within the tolerances of tests/test_gpu_parity.py.
    python tools/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lpfun_b200
from oracle import oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
t_end = time.time() + budget
n_cases = 0
while time.time() < t_end:
    m = int(rng.choice([1, 2, 2, 3, 3, 3, 4, 5, 6]))
    nmax = {1: 31, 2: 31, 3: 31, 4: 14, 5: 8, 6: 6}[m]
    n = int(rng.integers(1, nmax + 1))
    p = float(rng.choice([1.0, 2.0, 2.0, np.inf])) if m > 1 else 2.0
    if p == np.inf and m >= 3:
        n = min(n, {3: 14, 4: 7, 5: 4, 6: 3}[m])
    basis = str(rng.choice(["newton", "newton", "chebyshev"]))
    pre = bool(rng.integers(0, 2))
    colex = bool(rng.integers(0, 2))
    B = int(rng.choice([1, 3, 8, 33, 300]))
    kw = dict(lp_degree=p, basis=basis, precomputation=pre, colex_order=colex, report=False)
    try:
        t = lpfun_b200.Transform(m, n, exact_ifnt=True, **kw)
    except NotImplementedError:
        continue
    o = O.OracleTransform(m, n, p, precomputation=pre, colex_order=colex, basis=basis)
    N = len(t)
    if N * B > 4_000_000:
        B = max(1, 4_000_000 // N)
    v = rng.standard_normal((B, N))
    c = t.fnt(v)
    rows = sorted(set([0, B - 1, int(rng.integers(0, B))]))
    tag = f"m={m} n={n} p={p} {basis} pre={pre} colex={colex} B={B} N={N}"
    for b in rows:
        assert np.array_equal(c[b], o.fnt(v[b])), ("fnt", tag, b)
    r = t.ifnt(c)
    for b in rows:
        assert np.array_equal(r[b], o.ifnt(c[b])), ("ifnt exact", tag, b)
    tf = lpfun_b200.Transform(m, n, **kw)
    rf = tf.ifnt(c)
    assert np.max(np.abs(rf - r)) <= 1e-12 * max(1.0, np.max(np.abs(r))), ("ifnt fma", tag)
    i = int(rng.integers(0, m)); k = int(rng.integers(1, 4))
    try:
        d = t.dx(c[0], i, k)
        assert np.array_equal(d, o.dx(c[0], i, k)), ("dx", tag, i, k)
        dT = t.dxT(c[0], i, k)
        assert np.array_equal(dT, o.dxT(c[0], i, k)), ("dxT", tag, i, k)
    except NotImplementedError:
        pass
    P = int(rng.choice([1, 7, 300, 5000]))
    pts = rng.uniform(-1, 1, (P, m))
    cc = c[0] / (1.0 + np.arange(N)) ** 0.5
    want = o.eval(cc, pts)
    tol = (1e-12 if n < 25 else 1e-10) * max(1.0, np.max(np.abs(want)))
    for split in (0, -1):
        t.set_option("eval_split", split)
        got = t.eval(cc, pts)
        assert np.max(np.abs(got - want)) <= tol, ("eval", tag, P, split, float(np.max(np.abs(got - want))))
    t.close(); tf.close()
    n_cases += 1
    if n_cases % 10 == 0:
        print(f"{n_cases} cases ok, last: {tag}", flush=True)
print(f"FUZZ OK: {n_cases} random cases")
