"""Host-side logic of the multi-GPU path on CPU: row partition + final gather, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lpfun_b200.sharding import gather_rows, plan_push_chunks, row_block, sharded_apply


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


def test_plan_push_chunks_covers_rows_in_whole_rounds():
    for rows in (0, 1, 100, 148, 512, 2048, 2049):
        for sweep, copy in ((0.045, 0.00015), (0.057, 0.0012), (0.05, 0.0)):
            cuts = plan_push_chunks(rows, 148, sweep, copy)
            if rows == 0:
                assert cuts == []
                continue
            assert cuts[0][0] == 0 and cuts[-1][1] == rows and len(cuts) <= 5
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all(lo % 148 == 0 and hi > lo for lo, hi in cuts)
    # sweeps dominate: the last chunk (the only exposed copy) is the smallest; copies dominate: the first chunk is one round
    cuts = plan_push_chunks(2048, 148, 0.045, 0.00015)
    assert cuts[-1][1] - cuts[-1][0] == min(hi - lo for lo, hi in cuts) and len(cuts) >= 3
    cuts = plan_push_chunks(512, 148, 0.057, 0.0012)
    assert cuts[0] == (0, 148)


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


# ---- sharded SINGLE transform: slab / column layouts + all-to-all, CPU sweeps by the oracle ----
class _HostTransform:
    """What ShardedTransform reads from a Transform, built without a GPU from the oracle's matrices."""

    def __init__(self, m, n, p, precomputation, basis):
        from oracle import oracle as O

        self.o = O.OracleTransform(m, n, p, precomputation=precomputation, colex_order=False, basis=basis)
        self.spatial_dimension, self.polynomial_degree, self.device = m, n, 0
        self.multi_index_set = self.o.A
        mats = self.o.mats
        self._Vx_lt, self._inv_Vx_lt = mats["Vx_lt"], mats.get("inv_Vx_lt")
        self._Vx_ut, self._inv_Vx_ut = mats.get("Vx_ut"), mats.get("inv_Vx_ut")
        self._exact_ifnt = True


def _oracle_backend(A, packed, kind, solve, fma, dims, x):
    from oracle import oracle as O

    sw = O.Sweeper(A)
    n1 = int(round((np.sqrt(8 * len(packed) + 1) - 1) / 2))
    if kind == 0:
        return sw.lower_solve(packed, x, dims) if solve else sw.lower_matvec(packed, x, dims)
    assert not solve
    return sw.upper_matvec(packed, n1, x, dims)


def _sharded_worker(rank, world, port, cfg, q):
    from lpfun_b200.sharding import ShardedTransform

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, n, p, pre, basis = cfg
        ht = _HostTransform(m, n, p, pre, basis)
        st = ShardedTransform(ht, backend=_oracle_backend)
        N = len(ht.multi_index_set)
        rng = np.random.default_rng(3)
        v = rng.standard_normal(N)
        want_c = ht.o.fnt(v)
        want_v = ht.o.ifnt(want_c)
        c_cols = st.fnt(torch.as_tensor(st.slab_of(v).copy()))
        ok = np.array_equal(c_cols.numpy(), st.columns_of(want_c))
        v_slab = st.ifnt(c_cols)
        ok = ok and np.array_equal(v_slab.numpy(), st.slab_of(want_v))
        ok = ok and st.exchanges == (8 if basis == "chebyshev" else 4)  # fnt 1 (3), ifnt 3 (5) all-to-alls
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_slab_column_layout_is_a_partition():
    from lpfun_b200.sharding import SlabColumnLayout
    from oracle import oracle as O

    for m, n, p, G in ((2, 6, 2.0, 2), (3, 5, 2.0, 3), (3, 6, 1.0, 2), (4, 4, np.inf, 3), (3, 4, 2.0, 5), (3, 3, 2.0, 8)):
        A = O.lp_set(m, n, p)
        L = SlabColumnLayout(A, G)
        N = len(A)
        assert L.slab_elem[0] == 0 and L.slab_elem[-1] == N and int(L.col_elems.sum()) == N
        ids = np.arange(N, dtype=np.float64)
        seen = []
        for r in range(G):
            cols = L.full_to_columns(ids, r).astype(np.int64)
            S = L.column_set(r)
            assert len(S) == len(cols) and np.array_equal(A[cols, m - 1], S[:, 0])  # position k inside its column
            for j in np.unique(S[:, 1])[:4]:  # a column holds ONE base point
                assert len(np.unique(A[cols[S[:, 1] == j], : m - 1], axis=0)) == 1
            seen.append(cols)
            assert int(L.send_counts[r].sum()) == L.slab_elem[r + 1] - L.slab_elem[r]
            assert int(L.recv_counts[r].sum()) == L.col_elems[r]
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(N))
    with pytest.raises(ValueError):
        SlabColumnLayout(O.lp_set(1, 5, 2.0), 2)


@pytest.mark.parametrize("cfg", [(3, 6, 2.0, True, "newton"), (2, 7, 2.0, False, "newton"), (4, 4, 1.0, True, "newton"),
                                 (3, 5, 2.0, True, "chebyshev")])
def test_sharded_single_transform_world2_gloo(cfg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, cfg, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(0, True), (1, True)]
