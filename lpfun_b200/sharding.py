"""Row sharding of batched transforms across GPUs (one process per GPU).

The reference has no distributed code.  Batches of independent functions (rows of
the (B, N) array) and evaluation point sets shard trivially: every rank owns a contiguous block of
rows, runs the single-GPU path on it and -- only if the caller wants the full result everywhere --
a final all-gather collects the blocks (NCCL over NVLink on GPUs, gloo on CPU in the tests).
No collective sits on the transform path itself.
"""
from __future__ import annotations

from typing import Tuple

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def row_block(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous block; the first ``total % world`` ranks get one extra row."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("row_block: need world >= 1, 0 <= rank < world, total >= 0")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local, total: int, group=None):
    """All-gather row blocks produced with :func:`row_block` into the full (total, ...) tensor.

    Works for uneven blocks (pads to the largest block for the collective, trims afterwards)."""
    if dist is None or not dist.is_initialized():
        if local.shape[0] != total:
            raise ValueError("gather_rows without an initialised process group needs the full array")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, stop = row_block(total, rank, world)
    if local.shape[0] != stop - start:
        raise ValueError(f"rank {rank}: expected {stop - start} rows, got {local.shape[0]}")
    biggest = -(-total // world)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        a, b = row_block(total, r, world)
        pieces.append(out[r * biggest : r * biggest + (b - a)])
    return torch.cat(pieces, dim=0)


def sharded_apply(fn, full_rows, group=None, gather: bool = True):
    """Apply ``fn`` (e.g. ``t.fnt``) to this rank's block of ``full_rows`` and optionally all-gather."""
    if dist is None or not dist.is_initialized():
        return fn(full_rows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a, b = row_block(full_rows.shape[0], rank, world)
    local = fn(full_rows[a:b])
    return gather_rows(local, full_rows.shape[0], group) if gather else local
