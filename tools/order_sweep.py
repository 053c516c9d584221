#!/usr/bin/env python
"""k_whole2 item-order variants per coordinate (0 sorted, 1/2 classes of 4/8, 3 natural): device times (development helper)."""
import itertools, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lpfun_b200
t = lpfun_b200.Transform(3, 30, report=False, precompilation=False)
N = len(t); B = 4096
x = torch.randn(B, N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
def timeit(op):
    for _ in range(2): getattr(t, op)(x, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): getattr(t, op)(x, out=y)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
ref = {op: None for op in ("fnt", "ifnt")}
for op, key in (("fnt", "w2_order_exact"), ("ifnt", "w2_order_fma")):
    res = []
    for o0, o1, o2 in itertools.product(range(4), range(4), range(4)):
        t.set_option(key, o0 + 4 * o1 + 16 * o2)
        res.append((timeit(op), (o0, o1, o2)))
    res.sort()
    print(op, "best:", [(round(ms, 4), o) for ms, o in res[:6]], " sorted everywhere:", [round(ms, 4) for ms, o in res if o == (0, 0, 0)])
