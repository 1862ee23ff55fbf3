"""Batch sharding across the GPUs of one box (BASELINE config 5: N = 2^16, batch = 65536 over 8 x B200).

Transforms in a batch are independent -- the reference's batch is a plain serial loop over contiguous
chunks (src/array_utils.rs:164-170) and its users parallelise by giving each thread a slice
(examples/concurrency.rs:17-29).  So the batch is partitioned into contiguous ranges, one per rank
(`rustfft_b200.shard_range`), every rank runs its own plan replica on its own GPU, and NO collective
runs on the data path.  torch.distributed (NCCL over NVLink / NVSwitch) is used only when the batch
starts or must end on one rank: `scatter` / `gather` below move whole contiguous shards with
point-to-point sends inside one batched group.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import Fft, FftDirection, FftPlanner, shard_range


class ShardedFft:
    """One plan replica per rank + the shard bookkeeping.  Works with any backend of torch.distributed
    (NCCL on GPUs; gloo in the CPU tests, where the plan comes from the test-only emulation library)."""

    def __init__(self, planner: FftPlanner, n: int, direction: FftDirection = FftDirection.Forward, group=None):
        self.n = int(n)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.fft: Fft = planner.plan_fft(n, direction)

    def my_range(self, batch: int):
        return shard_range(batch, self.rank, self.world)

    def process_local(self, shard: torch.Tensor) -> torch.Tensor:
        """In place on this rank's shard (device tensor -> CUDA path, CPU tensor -> host-slice path)."""
        if shard.numel() == 0:
            return shard
        if shard.is_cuda:
            self.fft.process_device(shard)
        else:
            self.fft.process(shard.view(-1).numpy())
        return shard

    def scatter(self, full: Optional[torch.Tensor], batch: int, root: int = 0, device=None, dtype=None) -> torch.Tensor:
        """Root holds batch*n elements; every rank gets its contiguous shard."""
        lo, hi = self.my_range(batch)
        if self.world == 1:
            return full[lo * self.n: hi * self.n].clone()  # a copy, as for world > 1: process_local never touches `full`
        if self.rank == root:
            device, dtype = full.device, full.dtype
        elif device is None or dtype is None:
            raise ValueError("scatter(): ranks other than root must pass the shard's device and dtype")
        mine = torch.empty((hi - lo) * self.n, dtype=dtype, device=device)
        ops = []
        if self.rank == root:
            for r in range(self.world):
                a, b = shard_range(batch, r, self.world)
                if r == root:
                    mine.copy_(full[a * self.n: b * self.n])
                elif b > a:
                    ops.append(dist.P2POp(dist.isend, torch.view_as_real(full[a * self.n: b * self.n]), r, self.group))
        elif hi > lo:
            ops.append(dist.P2POp(dist.irecv, torch.view_as_real(mine), root, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return mine

    def gather(self, shard: torch.Tensor, batch: int, root: int = 0, out: Optional[torch.Tensor] = None):
        """Inverse of scatter: contiguous shards back to root, in batch order."""
        lo, hi = self.my_range(batch)
        if self.world == 1:
            return shard
        ops = []
        if self.rank == root:
            if out is None:
                out = torch.empty(batch * self.n, dtype=shard.dtype, device=shard.device)
            for r in range(self.world):
                a, b = shard_range(batch, r, self.world)
                if r == root:
                    out[a * self.n: b * self.n].copy_(shard)
                elif b > a:
                    ops.append(dist.P2POp(dist.irecv, torch.view_as_real(out[a * self.n: b * self.n]), r, self.group))
        elif hi > lo:
            ops.append(dist.P2POp(dist.isend, torch.view_as_real(shard), root, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out if self.rank == root else None
