#!/usr/bin/env python
"""Per-call latency of small single transforms: Python front end vs the bare C-ABI call (development helper)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lpfun_b200
from lpfun_b200 import _native as nat
for m, n, whole in ((2, 10, -1), (3, 20, -1), (3, 30, -1), (3, 30, 0), (6, 8, -1)):
    t = lpfun_b200.Transform(m, n, report=False, precompilation=False)
    t.set_option("use_whole", whole)
    N = len(t)
    x = torch.randn(N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    xh = np.random.rand(N)
    L = nat.lib()
    for _ in range(20): t.fnt(x, out=y)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): t.fnt(x, out=y)
    torch.cuda.synchronize(); t_py = (time.perf_counter() - t0) / 2000
    x2 = x.reshape(1, -1); y2 = y.reshape(1, -1)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): L.lpfnt_fnt(t._plan, x2.data_ptr(), y2.data_ptr(), 1, 1, s)
    torch.cuda.synchronize(); t_c = (time.perf_counter() - t0) / 2000
    t0 = time.perf_counter()
    for _ in range(500): t.fnt(xh)
    t_h = (time.perf_counter() - t0) / 500
    print(f"m={m} n={n} N={N} use_whole={whole}: fnt per call  device tensor {t_py*1e6:6.1f} us (bare C-ABI {t_c*1e6:6.1f} us)   NumPy in/out {t_h*1e6:6.1f} us")
    t.close()
