#!/usr/bin/env python
"""bench.py -- FNT+IFNT throughput of the B200 path, next to the CPU oracle port.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Contract (see DESIGN.md "Measurement"):
* one STEP = one FNT + one IFNT over one batch of synthetic smooth functions, device resident;
* headline workload = BASELINE config 4: Transform(3, 30), B = 4096 functions PER GPU
  (498 MB in, 498 MB out per transform: inputs are larger than the 126 MB L2, so no flush);
  the batch shards by rows with no data-path collective -> "scaling": "weak";
* `value` = 2*N*B*n_gpus / t_step / 1e9  GDOF/s, CUDA-event timed, max over ranks;
* also reported in the same JSON line: config 3 (d=6, n=20 single vector, L2 flushed between
  iterations; replicas only), config 5 (dx + eval), `roofline`, `cpu_baseline`, `e2e`, `clocks`.
* `--impl reference` times the CPU restatement of the reference algorithm (oracle/, "port") on the
  host cores with the same metric/config; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FNT+IFNT GDOF/s (FP64), batched d=3 n=30 (also d=6 n=20 single); % of HBM roofline"
C4 = dict(m=3, n=30)
C3 = dict(m=6, n=20)
C5 = dict(m=4, n=20)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# synthetic inputs, generated ON THE GRID the plan exposes
# ------------------------------------------------------------------------------------------
def sines_batch(grid: np.ndarray, B: int, seed: int, q: int = 4) -> np.ndarray:
    """B functions f_b = sum_q a sin(w.x + phi) sampled on grid -> (B, N) float64."""
    rng = np.random.default_rng(seed)
    m = grid.shape[1]
    out = np.empty((B, grid.shape[0]), dtype=np.float64)
    a = rng.uniform(0.5, 1.0, size=(B, q))
    w = rng.uniform(-3.0, 3.0, size=(B, q, m))
    ph = rng.uniform(0.0, 2 * np.pi, size=(B, q))
    step = 64
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        arg = np.einsum("nm,bqm->bqn", grid, w[b0:b1]) + ph[b0:b1, :, None]
        out[b0:b1] = np.einsum("bq,bqn->bn", a[b0:b1], np.sin(arg))
    return out


def smooth(grid: np.ndarray) -> np.ndarray:
    m = grid.shape[1]
    w = np.arange(1, m + 1, dtype=np.float64) / 6.0
    return np.exp(-np.sum(grid * grid, axis=1)) * np.sin(grid @ w)


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm, all host threads
# ------------------------------------------------------------------------------------------
def cpu_roundtrip_rate(m: int, n: int, budget_s: float, max_funcs: int, seed: int = 1234):
    """Time fnt+ifnt of the CPU oracle on as many functions as fit in budget_s.
    Returns (GDOF/s, functions timed, seconds, threads)."""
    from oracle import oracle as O

    O.use_all_cores()
    o = O.OracleTransform(m, n, 2.0)
    for h in range(m):
        o.sw.neighbors(h)  # plan build, not timed (the reference builds T/cs_T/e_T in its ctor too)
    vals = sines_batch(o.grid, min(max_funcs, 8), seed) if o.N < 100000 else smooth(o.grid)[None, :]
    o.ifnt(o.fnt(vals[0]))  # warm
    done, t0 = 0, time.perf_counter()
    while done < max_funcs:
        v = vals[done % len(vals)]
        o.ifnt(o.fnt(v))
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return 2.0 * o.N * done / dt / 1e9, done, dt, O.num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    rates, funcs = [], 0
    per_step_budget = 6.0
    total = args.steps + args.warmup
    tsum = 0.0
    for s in range(total):
        r, done, dt, thr = cpu_roundtrip_rate(C4["m"], C4["n"], per_step_budget, 4096, seed=1234 + s)
        if s >= args.warmup:
            rates.append(r); funcs = done; tsum += dt
    val = float(np.mean(rates))
    sample = f"{funcs} of {B} functions per step (fnt+ifnt each), d=3 n=30, oracle C port, {thr} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tsum / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batched fnt+ifnt, Transform(3,30), B={B} per GPU (CPU arm: bounded sample)", "m": 3, "n": 30, "N": 15216, "batch_per_gpu": B},
        "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
LABEL_NAMES = {0: "k_sweep_lines", 1: "k_sweep_groups", 2: "k_sweep_generic", 3: "k_eval_horner", 8: "k_permute", 11: "k_eval_pad", 12: "k_eval_block", 14: "k_whole", 16: "k_slab"}


def summarize_trace(recs):
    agg = {}
    for label, ms in recs:
        key = f"{LABEL_NAMES.get(label // 100, '?')}[h={label % 100}]"
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1; a[1] += ms
    total = sum(v[1] for v in agg.values()) or 1.0
    return {k: {"launches": v[0], "avg_ms": v[1] / v[0], "share": v[1] / total} for k, v in agg.items()}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import lpfun_b200
    from lpfun_b200 import _native as nat

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device -- lpfun_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_src = measured_peaks()
    B, K, W = args.batch, args.steps, args.warmup
    out = {}

    # ---------------- headline: batched C4 ----------------
    t = lpfun_b200.Transform(C4["m"], C4["n"], report=False, device=local)
    N = len(t)
    vals_h = sines_batch(t.grid, min(B, 256), seed=1234 + rank)
    reps = (B + len(vals_h) - 1) // len(vals_h)
    scale = (1.0 + 0.001 * np.arange(reps))[:, None, None]
    vals_full = (vals_h[None, :, :] * scale).reshape(-1, N)[:B]           # B distinct functions
    vals = torch.as_tensor(vals_full, device=dev)
    coef = torch.empty_like(vals)
    rec = torch.empty_like(vals)
    launches0 = t.launch_count
    for _ in range(W):
        t.fnt(vals, out=coef); t.ifnt(coef, out=rec)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l_before = t.launch_count
    barrier()
    e0.record()
    for _ in range(K):
        t.fnt(vals, out=coef); t.ifnt(coef, out=rec)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = t.launch_count - l_before
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / K
    value = 2.0 * N * B * world / (ms_step * 1e-3) / 1e9
    err = float((rec - vals).abs().max().item())
    # traced pass: per-kernel device times (CUDA events on the launching stream)
    t.set_option("trace", 1)
    for _ in range(min(K, 5)):
        t.fnt(vals, out=coef); t.ifnt(coef, out=rec)
    torch.cuda.synchronize()
    trace = summarize_trace(t.trace_read())
    t.set_option("trace", 0)
    dom = max(trace.items(), key=lambda kv: kv[1]["avg_ms"] * kv[1]["launches"]) if trace else (None, None)
    bytes_step = 32.0 * N * B
    achieved = bytes_step / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_step": bytes_step, "note": "16 B/DOF per transform (read+write once); whole step = all sweep launches",
                "kernels": trace, "dominant_kernel": dom[0]}
    if dom[0]:
        roofline["dominant_kernel_sweep_GBs"] = 16.0 * N * B / (dom[1]["avg_ms"] * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel, per launch, from the committed `ncu --set full` capture of this
        # very command (profiles/): only reported when kernel, N and batch match the capture
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
                tr = json.load(f)
            if dom[0].startswith(tr["kernel"]) and tr["N"] == N and tr["batch_per_gpu"] == B:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
                roofline["algorithmic_bytes_per_launch"] = 16.0 * N * B
        except (OSError, KeyError, ValueError):
            pass

    # ---------------- e2e: host (pinned) buffers through the public API ----------------
    Be = min(B, args.e2e_batch)
    hv = nat.pinned_empty((Be, N)); hc = nat.pinned_empty((Be, N)); hr = nat.pinned_empty((Be, N))
    hv[:] = vals_full[:Be]
    t.fnt(hv, out=hc); t.ifnt(hc, out=hr)
    barrier()
    Ke = max(1, min(K, 5))
    t0 = time.perf_counter()
    for _ in range(Ke):
        t.fnt(hv, out=hc); t.ifnt(hc, out=hr)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": 2.0 * N * Be * world * Ke / dt / 1e9, "unit": "GDOF/s",
           "h2d_bytes_per_step": int(2 * Be * N * 8), "d2h_bytes_per_step": int(2 * Be * N * 8),
           "batch": Be, "note": "t.fnt(host)->host, t.ifnt(host)->host; pinned NumPy buffers; wall clock incl. H2D/D2H"}
    assert np.max(np.abs(hr - hv)) < 1e-6
    for a in (hv, hc, hr):
        nat.pinned_free(a)
    del vals, coef, rec
    t.close()

    extras = {}
    if rank == 0 and not args.skip_extras:
        extras = run_extras(args, torch, lpfun_b200, nat, dev, local, hbm_peak)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        r, done, dtc, thr = cpu_roundtrip_rate(C4["m"], C4["n"], 12.0, 4096)
        cpu_baseline = {"value": r, "unit": "GDOF/s", "cores": thr, "kind": "port",
                        "sample": f"{done} of {B} functions (fnt+ifnt each) in {dtc:.1f} s, oracle C port, {thr} OpenMP threads, host has {os.cpu_count()} cpus"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"batched fnt+ifnt round trip, Transform(3,30) (N={N}), B={B} functions per GPU, colex_order=True, fnt bit-exact / ifnt FMA",
                       "m": 3, "n": 30, "N": N, "batch_per_gpu": B, "l2": "inputs (498 MB per array) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"rows sharded over {world} GPU(s), no data-path collective"},
            "roundtrip_max_abs_err": err,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches),
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, torch, lpfun_b200, nat, dev, local, hbm_peak):
    """Config 3 (d=6 n=20 single vector, L2 flushed between iterations), config 5 (dx + eval)."""
    ex = {}
    K = max(3, min(args.steps, 10))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # ---- config 3 ----
    t0 = time.time()
    t = lpfun_b200.Transform(C3["m"], C3["n"], report=False, device=local)
    build_s = time.time() - t0
    N = len(t)
    v = torch.as_tensor(smooth(t.grid), device=dev)
    c = torch.empty_like(v); r = torch.empty_like(v)
    for _ in range(3):
        t.fnt(v, out=c); t.ifnt(c, out=r)
    torch.cuda.synchronize()
    times = []
    for _ in range(K):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); t.fnt(v, out=c); t.ifnt(c, out=r); b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = float(np.median(times))
    t.set_option("trace", 1)
    for _ in range(3):
        flush.fill_(1)
        t.fnt(v, out=c); t.ifnt(c, out=r)
    torch.cuda.synchronize()
    tr = summarize_trace(t.trace_read())
    t.set_option("trace", 0)
    ach = 32.0 * N / (ms * 1e-3) / 1e9
    ex["d6_single"] = {"workload": "Transform(6,20) single vector fnt+ifnt, 256 MB L2 flush between iterations, replicas only",
                       "N": N, "ms_per_roundtrip": ms, "GDOF_s": 2.0 * N / (ms * 1e-3) / 1e9,
                       "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "kernels": tr},
                       "roundtrip_max_abs_err": float((r - v).abs().max().item()), "plan_build_s": build_s,
                       "plan_device_MB": t.device_bytes / 1e6}
    del v, c, r
    t.close()
    # ---- config 5: dx + eval ----
    t = lpfun_b200.Transform(C5["m"], C5["n"], report=False, device=local)
    N = len(t)
    cf = t.fnt(torch.as_tensor(smooth(t.grid), device=dev))
    d = torch.empty_like(cf)
    dx_us = []
    for i in range(4):
        for _ in range(3):
            t.dx(cf, i, 1, out=d)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            t.dx(cf, i, 1, out=d)
        b.record(); torch.cuda.synchronize()
        dx_us.append(a.elapsed_time(b) / 20 * 1e3)
    P = args.eval_points
    pts = torch.rand((P, 4), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1234)) * 2 - 1
    t.dx(cf, 0, 1, out=d)
    t.eval(d, pts[:100000])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); vals = t.eval(d, pts); b.record(); torch.cuda.synchronize()
    ems = a.elapsed_time(b)
    import ctypes as C
    f = C.c_double(0.0)
    nat.check(nat.lib().lpfnt_measure_fp64(local, 20000, C.byref(f)))
    pairs = float(P) * N / (ems * 1e-3)
    ex["c5_dx_eval"] = {"workload": f"Transform(4,20): dx(i,1) i=0..3; eval at {P} random points", "N": N,
                        "dx_us": dx_us, "eval_ms": ems, "eval_Gpair_s": pairs / 1e9,
                        "roofline": {"bound": "fp64", "achieved": 2 * pairs / 1e12, "peak": 2 * f.value / 1e12, "unit": "TFLOP/s",
                                     "frac": pairs / f.value, "peak_source": "lpfnt_measure_fp64 (dependent DFMA chains, this GPU, this run)"}}
    t.close()
    return ex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="functions per GPU")
    ap.add_argument("--e2e-batch", type=int, default=4096)
    ap.add_argument("--eval-points", type=int, default=10_000_000)
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
