#!/usr/bin/env python
"""Phase timeline of the whole-function kernel (k_whole2), CTA 0, first four functions (development helper)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lpfun_b200
from lpfun_b200 import _native as nat

t = lpfun_b200.Transform(3, 30, report=False, precompilation=False)
N = len(t)
x = torch.randn(1480, N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
L = nat.lib()
L.lpfnt_plan_timeline_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
NAMES = ["top", "data in", "S0 loaded", "S0 computed", "S0 stored", "S0 barrier", "S1 loaded", "S1 computed", "S1 stored", "S1 barrier",
         "S2 loaded+sync", "S2 computed", "S2 stored"]
for op in ("fnt", "ifnt"):
    for _ in range(2):
        getattr(t, op)(x, out=y)
    t.set_option("timeline", 1)
    getattr(t, op)(x, out=y)
    buf = np.zeros(4 * 24 * 16, dtype=np.int64)
    nat.check(L.lpfnt_plan_timeline_read(t._plan, buf.ctypes.data, buf.size))
    t.set_option("timeline", 0)
    st = buf.reshape(4, 24, 16)
    print(f"== {op}: cycles since the function's first stamp (function 2 of CTA 0); warp 0 = longest items, 22 = shortest")
    f = st[2]; t0 = f[:, 0].min()
    for w in (0, 1, 4, 8, 12, 16, 20, 22):
        print(f" warp {w:2d}: " + "  ".join(f"{NAMES[i]} {int(f[w, i] - t0):6d}" for i in range(13)))
    print("   per stamp: min / max over the 23 busy warps (argmax warp):")
    for i in range(13):
        col = f[:23, i] - t0
        print(f"     {NAMES[i]:16s} {int(col.min()):7d} {int(col.max()):7d}  (warp {int(col.argmax())})")
    print("   function period (top to top):", int(st[3][:, 0].min() - st[2][:, 0].min()), "cycles")
