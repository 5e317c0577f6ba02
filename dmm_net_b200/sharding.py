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


class FlatGradBucket:
    """ONE gradient bucket for the whole model: every parameter's ``.grad`` is a view into one flat fp32 buffer, so the
    data-parallel exchange of a training step is a single all-reduce (NCCL over NVLink on the GPU boxes) with no
    per-parameter hooks, no bucket copies and no per-parameter messages.  Reference: train.py:178-184 wraps the model in
    DDP (bucketed all-reduce) and THEN all-reduces every parameter tensor once more (train.py:62-68, one message per
    tensor, dividing before the async handle completed); the matching layer has no parameters, so this exchange is the
    only collective of the whole path.  Autograd accumulates into the views in place."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                                                        # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())

    def rendezvous(self):
        """4-byte all-reduce: returns (on the stream) once every rank has finished its backward -- separates the wait for the
        slowest rank from the time of the gradient exchange itself"""
        import torch.distributed as dist
        if not hasattr(self, "_probe"):
            self._probe = torch.zeros(1, device=self.flat.device)
        dist.all_reduce(self._probe)

    @property
    def nbytes(self):
        return self.flat.numel() * 4
