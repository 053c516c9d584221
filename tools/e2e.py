#!/usr/bin/env python
"""Host-pointer (NumPy in / NumPy out) throughput of fnt+ifnt for several chunk sizes (development helper)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpfun_b200
from lpfun_b200 import _native as nat
t = lpfun_b200.Transform(3, 30, report=False, precompilation=False)
N = len(t); B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hv = nat.pinned_empty((B, N)); hc = nat.pinned_empty((B, N)); hr = nat.pinned_empty((B, N))
hv[:] = np.random.default_rng(0).standard_normal((B, N))
for mb in (64, 32, 16, 8, 4, 2):
    t.set_option("host_chunk_bytes", mb << 20)
    t.fnt(hv, out=hc); t.ifnt(hc, out=hr)
    t0 = time.perf_counter()
    for _ in range(5):
        t.fnt(hv, out=hc); t.ifnt(hc, out=hr)
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {mb:3d} MB: {dt*1e3:7.3f} ms per fnt+ifnt  {2*N*B/dt/1e9:6.2f} GDOF/s   ({4*N*B*8/dt/1e9:5.1f} GB/s over PCIe, both directions)")
assert np.max(np.abs(hr - hv)) < 1e-6
