"""Host-side logic of the multi-GPU path: clips/problems are independent, so ranks shard them with no data-path
collective (reference: DistributedSampler(num_replicas, rank) in eval.py:57-59 and N independent eval processes,
scripts/eval/eval.sh:12-45).  Only reporting uses a collective: max over ranks of the device time."""
from __future__ import annotations

from typing import List

import torch


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin shard ``i % world == rank`` (no padding, no duplication: every clip is processed exactly once)."""
    assert 0 <= rank < world
    return list(range(rank, n_items, world))


def max_over_ranks(value: float, device=None) -> float:
    """max over ranks of a device-measured duration (identity when torch.distributed is not initialised)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.get_backend() == "nccl":
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.get_backend() == "nccl":
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(items_this_rank: int, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job items/s = (items of all ranks) / (max over ranks of the elapsed device time)."""
    total = sum_over_ranks(items_this_rank, device)
    worst = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (worst * 1e-3)
