"""Sharded single transform on GPUs, one process per GPU (torchrun or plain python for world 1):
every rank also runs the un-sharded transform and compares bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpfun_b200  # noqa: E402
from lpfun_b200.sharding import ShardedTransform  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    ok = True
    try:
        rng = np.random.default_rng(7)
        only = os.environ.get("SHARDED_CASE")
        cases = [dict(m=3, n=12), dict(m=4, n=9, lp_degree=1.0), dict(m=2, n=20, precomputation=False),
                 dict(m=3, n=8, basis="chebyshev"), dict(m=6, n=7), dict(m=3, n=30)]
        if os.environ.get("SHARDED_BIG"):
            cases = [dict(m=6, n=20)]  # BASELINE config 3 (N = 6 974 412), for the record: see DESIGN.md
        if only is not None:
            cases = [cases[int(only)]]
        for kw in cases:
            m, n = kw.pop("m"), kw.pop("n")
            t = lpfun_b200.Transform(m, n, colex_order=False, report=False, device=local, exact_ifnt=True, **kw)
            st = ShardedTransform(t)
            N = len(t)
            v = rng.standard_normal(N)
            want_c = t.fnt(v)
            want_v = t.ifnt(want_c)
            c_cols = st.fnt(torch.as_tensor(np.ascontiguousarray(st.slab_of(v)), device="cuda"))
            good = np.array_equal(c_cols.cpu().numpy(), st.columns_of(want_c))
            v_slab = st.ifnt(c_cols)
            good = good and np.array_equal(v_slab.cpu().numpy(), st.slab_of(want_v))
            # timing of the sharded round trip (device resident)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x = torch.as_tensor(np.ascontiguousarray(st.slab_of(v)), device="cuda")
            e0.record()
            for _ in range(5):
                st.ifnt(st.fnt(x))
            e1.record(); torch.cuda.synchronize()
            if rank == 0:
                print(f"sharded m={m} n={n} {kw} N={N} world={world}: bit-exact {good}, {e0.elapsed_time(e1) / 5:.3f} ms per fnt+ifnt, "
                      f"{st.exchanges} all-to-alls so far", flush=True)
            ok = ok and good
            st.close(); t.close()
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() == 1.0)
    finally:
        dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK" if ok else "SHARDED_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
