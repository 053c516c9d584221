#!/usr/bin/env python
"""Phase timeline of the whole-function kernel (k_whole), CTAs 0 and 1, first four functions of each (development helper).
Stamps per function: 0 top, 1 data in; per coordinate h: 2+4h fibre loaded (last coordinate: + barrier, prefetch issued),
3+4h triangle code done, 4+4h stored, 5+4h barrier passed (not after the last coordinate).
   python tools/timeline.py [n] [option=value ...]
(whole_split=1 needs a library built with LPFNT_TIMELINE_SPLIT=1: the stamped split kernel is not in the default build)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lpfun_b200
from lpfun_b200 import _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
t = lpfun_b200.Transform(3, n, report=False, precompilation=False)
for o in sys.argv[2:]:
    k, v = o.split("="); t.set_option(k, int(v))
N = len(t)
x = torch.randn(2960, N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
L = nat.lib()
S, W = 16, 24
NAMES = ["top", "data in"] + [f"S{h} {w}" for h in range(3) for w in ("loaded", "computed", "stored", "barrier")]
SPLIT = any(o.startswith("whole_split=1") for o in sys.argv[2:])
for op in ("fnt", "ifnt"):
    for _ in range(2):
        getattr(t, op)(x, out=y)
    t.set_option("timeline", 1)
    getattr(t, op)(x, out=y)
    buf = np.zeros(4 * 2 * W * S, dtype=np.int64)
    nat.check(L.lpfnt_plan_timeline_read(t._plan, buf.ctypes.data, buf.size))
    t.set_option("timeline", 0)
    if SPLIT:  # k_whole_split: 12 warps, 32 stamps per warp and function (13 per unit: P then Q)
        st = buf.reshape(4, 2, 12, 32)
        print(f"== {op} (n={n}, split kernel, 12 warps per CTA, two CTAs per SM): cycles since the CTA's first stamp of its function 2")
        for c in range(2):
            f = st[2, c]
            t0 = f[f[:, 0] > 0, 0].min()
            print(f" CTA {c}: period {int(st[3, c][st[3, c][:, 0] > 0, 0].min() - t0)} cycles (function 2 -> 3)")
            for w in (0, 1, 6, 11):
                for part, base in (("P", 0), ("Q", 13)):
                    print(f"   warp {w:2d} unit {part}: " + "  ".join(f"{NAMES[i]} {int(f[w, base + i] - t0)}" for i in range(13) if f[w, base + i] > 0))
        continue
    for nw in (24, 12):  # layout depends on the CTA size: [fn][cta][warps][stamps]
        st = buf[: 4 * 2 * nw * S].reshape(4, 2, nw, S)
        if st[2, 0, nw - 1, 0] > 0 or nw == 12:
            break
    print(f"== {op} (n={n}, {nw} warps per CTA): cycles since the CTA's first stamp of its function 2")
    for c in range(2):
        f = st[2, c]
        if f[:, 0].max() == 0:
            continue
        t0 = f[f[:, 0] > 0, 0].min()
        print(f" CTA {c}: period {int(st[3, c][st[3, c][:, 0] > 0, 0].min() - t0)} cycles (function 2 -> 3)")
        for w in sorted({0, 1, nw // 2, nw - 2, nw - 1}):
            print(f"   warp {w:2d}: " + "  ".join(f"{NAMES[i]} {int(f[w, i] - t0)}" for i in range(14) if f[w, i] > 0))
