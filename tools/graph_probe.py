import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch, lpfun_b200
t = lpfun_b200.Transform(4, 20, report=False)
N = len(t)
c = torch.randn(N, dtype=torch.float64, device="cuda")
d = [torch.empty_like(c) for _ in range(4)]
for i in range(4): t.dx(c, i, 1, out=d[i])
torch.cuda.synchronize()
ref = [x.clone() for x in d]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(4): t.dx(c, i, 1, out=d[i])
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    for i in range(4): t.dx(c, i, 1, out=d[i])
for x in d: x.zero_()
g.replay(); torch.cuda.synchronize()
print("graph replay equals eager:", all(torch.equal(a, b) for a, b in zip(d, ref)))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): g.replay()
b.record(); torch.cuda.synchronize()
print(f"graph: {a.elapsed_time(b)/50/4*1e3:.1f} us per dx")
a.record()
for _ in range(50):
    for i in range(4): t.dx(c, i, 1, out=d[i])
b.record(); torch.cuda.synchronize()
print(f"eager: {a.elapsed_time(b)/50/4*1e3:.1f} us per dx")
# fnt+ifnt single vector d6? keep small: d=3 n=12 batch 16
t2 = lpfun_b200.Transform(3, 12, report=False)
v = torch.randn(16, len(t2), dtype=torch.float64, device="cuda"); cc = torch.empty_like(v); r = torch.empty_like(v)
t2.fnt(v, out=cc); t2.ifnt(cc, out=r); torch.cuda.synchronize()
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    t2.fnt(v, out=cc); t2.ifnt(cc, out=r)
cc0 = cc.clone(); cc.zero_(); g2.replay(); torch.cuda.synchronize()
print("fnt graph equals eager:", torch.equal(cc, cc0))
