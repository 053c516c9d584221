"""Fused all-gather through the whole-function kernel's stores (lpfun_b200.sharding.PushGather), one process per GPU
(torchrun, or plain python for a world of one): every rank compares the gathered batch with its own un-sharded result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpfun_b200  # noqa: E402
from lpfun_b200.sharding import PushGather, row_block  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29547")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    ok = True
    try:
        for (m, n, total, exact) in ((3, 30, 64 * world + 3, True), (3, 30, 600, False), (3, 20, 96, True)):
            t = lpfun_b200.Transform(m, n, report=False, device=local, exact_ifnt=exact)
            t.set_option("use_whole", 1)
            N = len(t)
            rng = np.random.default_rng(11)  # the same batch on every rank
            v = torch.as_tensor(rng.standard_normal((total, N)), device="cuda")
            want_c = t.fnt(v)
            want_r = t.ifnt(want_c)
            a, b = row_block(total, rank, world)
            for use_mc in (True, False):
                pg = PushGather(t, total, use_multicast=use_mc)
                pg.full.zero_()
                torch.cuda.synchronize(); dist.barrier()
                t.fnt(v[a:b], out=pg.local_rows())      # exact: bit-identical everywhere
                pg.barrier()
                torch.cuda.synchronize()
                good = bool(torch.equal(pg.full, want_c))
                c_local = pg.full[a:b].clone()
                torch.cuda.synchronize(); dist.barrier()
                t.ifnt(c_local, out=pg.local_rows())
                pg.barrier()
                torch.cuda.synchronize()
                good = good and bool(torch.equal(pg.full, want_r))
                # timing: push against the plain kernel + NCCL all-gather
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier(); torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    t.ifnt(c_local, out=pg.local_rows()); pg.barrier()
                e1.record(); torch.cuda.synchronize()
                ms_push = e0.elapsed_time(e1) / 5
                pg.close()
                # the streaming copy kernel: rows written locally, then pushed in two pieces
                pc = PushGather(t, total, use_multicast=use_mc, fused=False)
                pc.full.zero_()
                torch.cuda.synchronize(); dist.barrier()
                t.ifnt(c_local, out=pc.local_rows())
                half = (b - a) // 2
                pc.push_rows(0, half, engine="ce" if use_mc else "sm"); pc.push_rows(half, b - a)
                pc.barrier()
                torch.cuda.synchronize()
                good = good and bool(torch.equal(pc.full, want_r))
                pc.close()
                if rank == 0:
                    print(f"push m={m} n={n} total={total} world={world} multicast={pg.multicast} (asked {use_mc}): equal {good}, {ms_push:.3f} ms per ifnt+gather", flush=True)
                ok = ok and good
                torch.cuda.synchronize(); dist.barrier()
            t.close()
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() == 1.0)
    finally:
        dist.destroy_process_group()
    if rank == 0:
        print("PUSH_OK" if ok else "PUSH_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
