#!/usr/bin/env python
"""Quick device-resident timing of one config (development helper, not the bench contract).
   python tools/quick.py --m 3 --n 30 --batch 1024 [--opt whole_split=1] [--opt whole_skew=0]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lpfun_b200

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=3); ap.add_argument("--n", type=int, default=30)
ap.add_argument("--batch", type=int, default=1024); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--opt", action="append", default=[]); ap.add_argument("--flush", action="store_true")
ap.add_argument("--colex", type=int, default=1); ap.add_argument("--trace", type=int, default=1)
ap.add_argument("--ops", default="fnt,ifnt")
ap.add_argument("--sweep", default="", help="option=v1,v2,...: time the step for every value")
ap.add_argument("--check", type=int, default=1, help="compare against the per-coordinate generic-path kernels (bit-exact for fnt)")
a = ap.parse_args()
t0 = time.time()
t = lpfun_b200.Transform(a.m, a.n, report=False, colex_order=bool(a.colex), precompilation=False)
N = len(t)
for o in a.opt:
    k, v = o.split("="); t.set_option(k, int(v))
x = torch.randn(a.batch, N, dtype=torch.float64, device="cuda")
y = torch.empty_like(x); z = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if a.flush else None
ops = a.ops.split(",")
def step():
    if "fnt" in ops: t.fnt(x, out=y)
    if "ifnt" in ops: t.ifnt(y, out=z)
    if "dx" in ops: t.dx(y, a.m - 1, 1, out=z)
for _ in range(3): step()
torch.cuda.synchronize()
if a.check and 'fnt' in ops:
    nb = min(a.batch, 8)
    got = t.fnt(x[:nb].clone()); gotr = t.ifnt(got)
    t.set_option('use_whole', 0)
    ref = t.fnt(x[:nb].clone()); refr = t.ifnt(ref)
    t.set_option('use_whole', -1)
    for o in a.opt:
        k_, v_ = o.split('='); t.set_option(k_, int(v_))
    print('   check vs line/fibre kernels: fnt bit-equal', bool(torch.equal(got, ref)), ' ifnt max rel', float(((gotr-refr).abs().max()/refr.abs().max()).item()))
def timed():
    ts = []
    for _ in range(a.iters):
        if flush is not None: flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
if a.sweep:
    key, vals = a.sweep.split('=')
    for v in vals.split(','):
        t.set_option(key, int(v)); step(); torch.cuda.synchronize()
        per = []
        for op in ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters): (t.fnt(x, out=y) if op == 'fnt' else t.ifnt(y, out=z))
            e1.record(); torch.cuda.synchronize(); per.append(e0.elapsed_time(e1) / a.iters)
        print(f'   {key}={v}: ' + '  '.join(f'{o} {p_:.4f} ms' for o, p_ in zip(ops, per)) + f'  step {sum(per):.4f} ms')
    sys.exit(0)
ts = []
for _ in range(a.iters):
    if flush is not None: flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
nops = sum(o in ops for o in ("fnt", "ifnt", "dx"))
print(f"m={a.m} n={a.n} N={N} B={a.batch} opts={a.opt} ops={ops}: {ms:.4f} ms/step  {nops*N*a.batch/ms/1e6:.1f} GDOF/s  {16*nops*N*a.batch/ms/1e6:.0f} GB/s (16B/DOF)  build {time.time()-t0:.1f}s")
if a.trace:
    t.set_option("trace", 1); step(); torch.cuda.synchronize()
    agg = {}
    for lab, m_ in t.trace_read():
        g = agg.setdefault(lab, [0, 0.0]); g[0] += 1; g[1] += m_
    for lab, (c, s) in sorted(agg.items()): print(f"   label {lab}: {c} launches, total {s:.4f} ms, avg {s/c*1e3:.1f} us")
