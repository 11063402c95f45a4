"""Data-parallel mapping: one process per GPU, mapping rays sharded across ranks, ONE
all-reduce over a flat gradient bucket per iteration (NCCL over NVLink / NVSwitch; gloo on
CPU for the host-logic tests).  The reference has no multi-GPU path at all (SURVEY section 2,
"Parallelism strategies"); this is the new capability north_star asks for.

Parameters are replicated; every rank renders its own slice of the ray batch; the kernels'
batch-global quantities (Co-SLAM: n_fs / n_sdf / n_valid and R in the loss means; NICE:
max(target_d)) are made global with tiny integer / scalar all-reduces so that the summed
gradient equals the single-GPU gradient of the whole batch; Adam then runs redundantly and
bit-identically on every rank.  Tracking stays on one GPU (per-frame sequential).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


class MappingDataParallel:
    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = list(params)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # ---- sharding -----------------------------------------------------------
    def shard(self, n: int) -> slice:
        """Contiguous slice [rank*n/W, (rank+1)*n/W) of a batch of n rays."""
        lo = (n * self.rank) // self.world
        hi = (n * (self.rank + 1)) // self.world
        return slice(lo, hi)

    def broadcast_params(self, src: int = 0):
        if self.world > 1:
            for p in self.params:
                dist.broadcast(p.data, src, group=self.group)

    # ---- collectives --------------------------------------------------------
    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_grads(self, extra: Optional[List[torch.Tensor]] = None):
        """Sum the gradients of all replicated parameters (+ `extra` tensors, e.g. the loss
        scalars) with ONE collective over a flat bucket."""
        if self.world == 1:
            return
        tensors = [p.grad for p in self.params if p.grad is not None]
        if extra:
            tensors = tensors + list(extra)
        flat = _flatten_dense_tensors(tensors)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        for t, f in zip(tensors, _unflatten_dense_tensors(flat, tensors)):
            t.copy_(f)
