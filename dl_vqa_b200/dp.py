"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce of bucketed gradients over
NVLink/NVSwitch, overlapped with the rest of backward.

The reference is single-GPU (main.py:23); this layer is new.  VqaNet's backward finishes parameter groups
in a fixed order (classifier 22.8 MB -> attention 9.4 MB -> question encoder 61 MB -> image encoder 1.5 MB)
and fires `grad_ready_hook` after each group; every group becomes one bucket that is all-reduced
asynchronously on NCCL's stream while the remaining (conv-dominated) backward runs.  `finish()` makes the
compute stream wait for the reductions and writes the averaged gradients back in place.
The path shards by samples only (no cross-sample statistic anywhere in models/model.py), so the single
exchange step is this all-reduce; no other collective exists.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class GradientAllReduce:
    def __init__(self, model, process_group=None, average: bool = True):
        self.model = model
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        self._pending: List[tuple] = []
        if self.world > 1:
            model.grad_ready_hook = self._on_group_ready

    def _on_group_ready(self, named_grads):
        grads = [g for _, g in named_grads]
        flat = torch.cat([g.reshape(-1) for g in grads])           # one bucket per finished stage
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self._pending.append((work, flat, grads))

    def finish(self):
        """Call after loss.backward(): waits (stream-wise) for the buckets and scatters them back."""
        for work, flat, grads in self._pending:
            work.wait()
            if self.average:
                flat.div_(self.world)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        self._pending.clear()

    def broadcast_parameters(self, src: int = 0):
        if self.world > 1:
            for p in self.model.parameters():
                dist.broadcast(p.data, src=src, group=self.pg)
