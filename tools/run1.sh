set -x
python -m pytest tests -m gpu -x -q -k "pageable or golden or device_tensor" 2>&1 | tail -3
python - <<'PY'
import numpy as np, time, lpfun_b200
t = lpfun_b200.Transform(3, 30, report=False)
N = len(t)
for B in (1024, 4096):
    v = np.random.default_rng(0).standard_normal((B, N))
    for thr in (0, 1, 2, 4, 6, 8):
        t.set_option("host_threads", thr)
        r = t.ifnt(t.fnt(v))
        t0 = time.perf_counter()
        for _ in range(2): r = t.ifnt(t.fnt(v))
        dt = (time.perf_counter() - t0) / 2
        print(f"pageable B={B} host_threads={thr}: {dt*1e3:.1f} ms per round trip = {2*N*B/dt/1e9:.2f} GDOF/s", flush=True)
    out = np.empty_like(v); out2 = np.empty_like(v)
    t.set_option("host_threads", 0)
    t.fnt(v, out=out); t.ifnt(out, out=out2)
    t0 = time.perf_counter()
    for _ in range(2): t.fnt(v, out=out); t.ifnt(out, out=out2)
    dt = (time.perf_counter() - t0) / 2
    print(f"pageable B={B} preallocated outputs: {dt*1e3:.1f} ms = {2*N*B/dt/1e9:.2f} GDOF/s", flush=True)
PY
