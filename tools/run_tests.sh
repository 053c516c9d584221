set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import torch, time, numpy as np, lpfun_b200
t = lpfun_b200.Transform(6, 20, report=False, precompilation=False)
c = torch.randn(len(t), dtype=torch.float64, device="cuda")
for P in (1, 10, 100, 1000, 8000):
    pts = torch.rand((P, 6), dtype=torch.float64, device="cuda") * 2 - 1
    for split in (1, 0):
        if split == 0 and P > 10: continue
        t.set_option("eval_split", split)
        t.eval(c, pts); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): v = t.eval(c, pts)
        torch.cuda.synchronize()
        print(f"d6 n20 eval of {P} points, split={split}: {(time.perf_counter()-t0)/3*1e3:.3f} ms", flush=True)
print("plan MB", t.device_bytes/1e6)
PY
