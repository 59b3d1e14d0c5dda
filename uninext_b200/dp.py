"""Data-parallel plumbing for the hot path: frames are sharded across ranks (one process per GPU) and parameter
gradients are combined with ONE all-reduce per step over a flat bucket.

Reference behaviour being replaced: detectron2's DDP wrap (detectron2/engine/defaults.py:60-79,380-381) around a model
fed by a rank-strided sampler (detectron2/data/samplers/distributed_sampler.py:60).  The op itself is rank-local
(SURVEY.md section 8e): nothing is exchanged on the data path, only gradients after backward.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_frames(n_frames: int, world: int, rank: int) -> range:
    """Rank-strided assignment, like the reference's TrainingSampler: frame i goes to rank i % world."""
    return range(rank, n_frames, world)


class FlatGradBucket:
    """All parameter gradients live in one contiguous buffer (each ``p.grad`` is a view into it), so the step's
    gradient exchange is a single NCCL all-reduce over NVLink instead of one per tensor/bucket."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("FlatGradBucket needs parameters of one device and dtype")
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None, async_op: bool = False):
        """Sum over ranks, divide by world size (DDP semantics). No-op for world size 1."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        self.flat.div_(world)                       # pre-divide: keeps the sum in range for low-precision buckets
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()
