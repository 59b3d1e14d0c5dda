"""Data-parallel plumbing for the hot path: frames are sharded across ranks (one process per GPU) and parameter
gradients are combined over ONE flat buffer.

Reference behaviour being replaced: detectron2's DDP wrap (detectron2/engine/defaults.py:60-79,380-381) around a model
fed by a rank-strided sampler (detectron2/data/samplers/distributed_sampler.py:60).  The op itself is rank-local
(SURVEY.md section 8e): nothing is exchanged on the data path, only gradients.

Two ways to run the exchange:
  * ``all_reduce_mean()``            -- one all-reduce of the whole buffer after backward (also what follows a CUDA-graph
                                       replay of the step);
  * ``overlap=True`` + ``finish()``  -- like DDP, the buffer is cut into contiguous slices and each slice is all-reduced
                                       asynchronously from a post-accumulate-grad hook as soon as backward has produced
                                       every gradient in it, so the exchange over NVLink runs under the rest of backward.
                                       Slices are launched in one fixed order (last slice first) on every rank.

Every ``p.grad`` is a view into the flat buffer.  Anything that replaces ``p.grad`` breaks that aliasing --
``optimizer.zero_grad()`` / ``module.zero_grad()`` default to ``set_to_none=True`` and do exactly that.  Use
``bucket.zero_()`` (or ``zero_grad(set_to_none=False)``); ``all_reduce_mean()`` / ``finish()`` verify the aliasing and
repair it (copying a stray gradient back into the buffer) instead of silently reducing stale zeros.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_frames(n_frames: int, world: int, rank: int) -> range:
    """Rank-strided assignment, like the reference's TrainingSampler: frame i goes to rank i % world."""
    return range(rank, n_frames, world)


def _distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class FlatGradBucket:
    def __init__(self, params: Iterable[torch.nn.Parameter], overlap: bool = False, slice_bytes: int = 8 << 20,
                 group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("FlatGradBucket needs parameters of one device and dtype")
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        self.group = group
        self._views, self._offsets = [], []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self._views.append(v)
            self._offsets.append(off)
            off += p.numel()
        # contiguous slices of ~slice_bytes, cut at parameter boundaries
        self._slices, self._slice_of = [], []          # (start, end, first param, last param + 1); param -> slice
        per = max(1, slice_bytes // self.flat.element_size())
        start_p = 0
        for i, p in enumerate(self.params):
            end = self._offsets[i] + p.numel()
            self._slice_of.append(len(self._slices))
            if end - self._offsets[start_p] >= per or i == len(self.params) - 1:
                self._slices.append((self._offsets[start_p], end, start_p, i + 1))
                start_p = i + 1
        self.overlap = bool(overlap)
        self._pending = [0] * len(self._slices)
        self._works = []
        self._next = len(self._slices) - 1               # slices are launched last-to-first (backward order)
        self._hooks = []
        if self.overlap:                    # hooks stay registered; `self.overlap = False` silences them (e.g. while a
            for i, p in enumerate(self.params):     # CUDA graph of the step is captured: collectives stay outside the graph)
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
        self._arm()

    # ---- aliasing guard ------------------------------------------------------------------------------------------
    def check_views(self) -> int:
        """Re-bind any ``p.grad`` that no longer aliases the flat buffer (e.g. after ``zero_grad(set_to_none=True)``),
        copying its values in.  Returns how many parameters had to be repaired."""
        fixed = 0
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is None:
                v.zero_()
                p.grad = v
                fixed += 1
            elif g.data_ptr() != v.data_ptr() or g.shape != v.shape or g.dtype != v.dtype:
                v.copy_(g)
                p.grad = v
                fixed += 1
        return fixed

    # ---- plain path ----------------------------------------------------------------------------------------------
    def zero_(self):
        self.flat.zero_()
        self._arm()

    def all_reduce_mean(self, group=None, async_op: bool = False):
        """Sum over ranks, divide by world size (DDP semantics). No-op for world size 1."""
        group = group if group is not None else self.group
        self.check_views()
        if not _distributed(group):
            return None
        world = dist.get_world_size(group)
        self.flat.div_(world)                       # pre-divide: keeps the sum in range for low-precision buckets
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    # ---- overlapped path -----------------------------------------------------------------------------------------
    def _arm(self):
        self._pending = [e - s for (_, _, s, e) in self._slices]      # gradients still missing per slice
        self._works = []
        self._next = len(self._slices) - 1

    def _make_hook(self, i):
        def hook(p):
            if not self.overlap:
                return
            v = self._views[i]
            if p.grad is not v and (p.grad is None or p.grad.data_ptr() != v.data_ptr()):
                v.copy_(p.grad)                      # the aliasing was broken since the last step: repair in place
                p.grad = v
            k = self._slice_of[i]
            if self._next < 0 and _distributed(self.group):
                raise RuntimeError("FlatGradBucket(overlap=True): a gradient arrived after every slice of this step had been "
                                   "all-reduced -- call zero_() (which re-arms the exchange) before each backward; gradient "
                                   "accumulation over several backward passes needs overlap=False")
            self._pending[k] -= 1
            self._launch_ready()
        return hook

    def _launch_ready(self, force: bool = False):
        if not _distributed(self.group):
            return
        world = dist.get_world_size(self.group)
        while self._next >= 0 and (force or self._pending[self._next] <= 0):
            s, e, _, _ = self._slices[self._next]
            part = self.flat[s:e]
            part.div_(world)
            self._works.append(dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._next -= 1

    def finish(self):
        """End of backward in overlap mode: launch whatever is still unlaunched (parameters that received no gradient
        this step keep their slice zero) and make the current stream wait for every slice."""
        if not self.overlap:
            return self.all_reduce_mean()
        self.check_views()
        self._launch_ready(force=True)
        for w in self._works:
            w.wait()
        self._works = []
        return None

    @property
    def n_slices(self) -> int:
        return len(self._slices)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()
