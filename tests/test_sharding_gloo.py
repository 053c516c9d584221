"""Host-side logic of the multi-GPU path on CPU: row partition + final gather, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lpfun_b200.sharding import gather_rows, row_block, sharded_apply


def test_row_block_partitions():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [row_block(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_block(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 5, dtype=torch.float64).reshape(total, 5)
        # stand-in for t.fnt on the local block: any row-wise map
        out = sharded_apply(lambda x: x * 2.0 + 1.0, full)
        ok = bool(torch.equal(out, full * 2.0 + 1.0))
        a, b = row_block(total, rank, world)
        local = sharded_apply(lambda x: x * 2.0 + 1.0, full, gather=False)
        ok = ok and tuple(local.shape) == (b - a, 5)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == world
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_rows_world2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]
