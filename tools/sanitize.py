#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): the kernels with hand-rolled synchronisation
(k_whole: bulk copies + mbarrier + in-place shared-memory sweeps; k_slab: cp.async staging) on shapes that make every CTA
loop over several functions, plus the per-coordinate kernels and eval.  Checks the results against the fixtures.
    compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import lpfun_b200
from conftest import load_case, case_kwargs

def make_transform(T, g, **kw):
    kwargs = case_kwargs(g)
    kwargs.update(report=False, precompilation=False)
    kwargs.update(kw)
    return T(**kwargs)

def check(name, opts, reps):
    g = load_case(name)
    t = make_transform(lpfun_b200.Transform, g, exact_ifnt=True)
    for k, v in opts.items():
        t.set_option(k, v)
    got = t.fnt(np.tile(g["values"], (reps, 1)))
    assert np.array_equal(got, np.tile(g["fnt"], (reps, 1))), (name, opts)
    back = t.ifnt(np.tile(g["fnt"], (reps, 1)))
    assert np.array_equal(back, np.tile(g["ifnt"], (reps, 1))), (name, opts)
    tf = make_transform(lpfun_b200.Transform, g)
    for k, v in opts.items():
        tf.set_option(k, v)
    r = tf.ifnt(np.tile(g["fnt"], (reps, 1)))
    assert np.max(np.abs(r - np.tile(g["ifnt"], (reps, 1)))) <= 1e-12 * np.max(np.abs(g["ifnt"]))
    m = int(g["m"])
    assert np.array_equal(t.dx(g["fnt"][0], m - 1, 1), g[f"dx_{m-1}_1"])
    t.eval(g["fnt"][0], g["points"])
    t.close(); tf.close()
    print("ok", name, opts, flush=True)

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 160
check("d3n12_p1", {"use_whole": 1}, reps)                                   # k_whole, odd N, two CTAs per SM
check("d3n12_p1", {"use_whole": 1, "whole_fibres": 2}, reps)                # two fibres side by side
check("c2_d3n20_leja", {"use_whole": 1}, reps)                              # colex gather / scatter
check("c2_d3n20_leja", {"use_whole": 0, "use_slab": 2}, 8)                  # k_slab + k_sweep_groups
check("d4n8", {"use_whole": 0}, 4)
print("SANITIZE_WORKLOAD_OK")
