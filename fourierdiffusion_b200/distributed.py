"""Multi-GPU plumbing of the sampler: one process per GPU, series sharded by global index, ONE all-gather at the end.

Every series is independent (the reference's batch loop, sampler.py:63-107, carries no state across series), so ranks
exchange nothing while sampling; the noise of a series is keyed by its GLOBAL index (csrc/fd_philox.cuh), so 1/2/4/8-GPU
runs return the same samples.  The only collective is the all-gather of the finished (n_local, L, C) fp32 shards
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of global series indices owned by `rank`; sizes differ by at most one."""
    assert 0 <= rank < world_size and n >= 0
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world_size: int) -> List[int]:
    return [shard_range(n, r, world_size)[1] - shard_range(n, r, world_size)[0] for r in range(world_size)]


def all_gather_series(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Gather the ranks' (n_local, L, C) shards into the full (n_total, L, C) tensor on every rank, in global index order.

    Uses a single all_gather_into_tensor when the shards are equal-sized (the normal case: batch divisible by the GPU
    count), else one padded all_gather followed by trimming.  `local` lives on the device of the process group's backend.
    """
    rank, ws = world()
    if ws == 1:
        assert local.shape[0] == n_total
        return local
    sizes = shard_sizes(n_total, ws)
    assert local.shape[0] == sizes[rank], f"rank {rank} holds {local.shape[0]} series, expected {sizes[rank]}"
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if len(set(sizes)) == 1:
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out
    m = max(sizes)
    padded = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((ws * m,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    return torch.cat([buf[r * m : r * m + sizes[r]] for r in range(ws)], dim=0)
