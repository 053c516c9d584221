import sys, os, time, numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpfun_b200
from lpfun_b200 import _native as nat
m, n, P = int(sys.argv[1]), int(sys.argv[2]), int(float(sys.argv[3]))
basis = sys.argv[4] if len(sys.argv) > 4 else 'newton'
t = lpfun_b200.Transform(m, n, report=False, precompilation=False, basis=basis)
N = len(t)
c = torch.randn(N, dtype=torch.float64, device="cuda")
pts = torch.rand((P, m), dtype=torch.float64, device="cuda") * 2 - 1
t.eval(c, pts[:1000]); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); v = t.eval(c, pts); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
f = C.c_double(0.0); nat.check(nat.lib().lpfnt_measure_fp64(0, 20000, C.byref(f)))
print(f"eval [{basis}] m={m} n={n} N={N} P={P}: {ms:.2f} ms  {P*N/ms/1e6:.1f} Gpair/s = {100*P*N/(ms*1e-3)/f.value:.1f} % of measured FP64 FMA peak ({f.value/1e12:.2f} TFMA/s)")
