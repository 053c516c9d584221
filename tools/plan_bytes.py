"""Device memory held by a plan at the stages of a session (creation, first transforms, first evals)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lpfun_b200

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6, 20)
t = lpfun_b200.Transform(m, n, report=False)
N = len(t)
print(f"Transform({m},{n}) N={N}: after creation {t.device_bytes/1e6:.1f} MB")
v = torch.rand(N, dtype=torch.float64, device="cuda")
c = t.fnt(v); r = t.ifnt(c); torch.cuda.synchronize()
print(f"  after fnt+ifnt {t.device_bytes/1e6:.1f} MB")
d = t.dx(c, 0, 1); torch.cuda.synchronize()
print(f"  after dx {t.device_bytes/1e6:.1f} MB")
p = torch.rand((10, m), dtype=torch.float64, device="cuda") * 2 - 1
t.eval(c, p); torch.cuda.synchronize()
print(f"  after eval(10 points) {t.device_bytes/1e6:.1f} MB")
p = torch.rand((100000, m), dtype=torch.float64, device="cuda") * 2 - 1
t.eval(c, p); torch.cuda.synchronize()
print(f"  after eval(1e5 points) {t.device_bytes/1e6:.1f} MB")
