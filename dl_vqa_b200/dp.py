"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce of bucketed gradients over
NVLink/NVSwitch, overlapped with the rest of backward.

The reference is single-GPU (main.py:23); this layer is new.  VqaNet's backward finishes parameter groups
in a fixed order (classifier 22.8 MB -> attention 9.4 MB -> question encoder 61 MB -> image encoder 1.5 MB)
and fires `grad_ready_hook` after each group; every group becomes one bucket that is all-reduced
asynchronously on NCCL's stream while the remaining (conv-dominated) backward runs.  `finish()` makes the
compute stream wait for the reductions and writes the averaged gradients back in place.
With the model's gradient arena (`VqaNet.use_gradient_arena`, switched on by this wrapper when available) a stage's
gradients ARE one flat buffer: the all-reduce runs in place on it, there is no concatenate / scatter-back copy, and
the division by the world size is folded into the optimizer (`FusedAdam.step(grad_scale=ddp.grad_scale)`).
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
        self.in_place = False
        if self.world > 1:
            model.grad_ready_hook = self._on_group_ready
            if hasattr(model, "use_gradient_arena"):
                model.use_gradient_arena(True)
                self.in_place = True

    @property
    def grad_scale(self) -> float:
        """Factor the optimizer must apply to the (summed) gradients: 1/world in in-place mode with averaging,
        else 1 (finish() has already divided)."""
        return 1.0 / self.world if (self.in_place and self.average and self.world > 1) else 1.0

    def _bucket_of(self, named_grads):
        """The arena bucket that holds exactly these gradients, or None."""
        buckets = self.model.gradient_buckets() if hasattr(self.model, "gradient_buckets") else None
        if not buckets or not named_grads:
            return None
        stage = named_grads[0][0].split(".")[0]
        flat = buckets.get(stage)
        if flat is None:
            return None
        base, end = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        for n, g in named_grads:
            if not n.startswith(stage + ".") or not (base <= g.data_ptr() < end):
                return None
        return flat

    def _on_group_ready(self, named_grads):
        flat = self._bucket_of(named_grads) if self.in_place else None
        if flat is not None:                                          # all-reduce the stage's bucket in place
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._pending.append((work, None, None))
            return
        self.in_place = False                                         # fresh tensors this step: classic path
        grads = [g for _, g in named_grads]
        flat = torch.cat([g.reshape(-1) for g in grads])           # one bucket per finished stage
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self._pending.append((work, flat, grads))

    def finish(self):
        """Call after loss.backward(): waits (stream-wise) for the buckets and scatters them back."""
        for work, flat, grads in self._pending:
            work.wait()
            if flat is None:
                continue                                              # in place: nothing to scatter, scale in the optimizer
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
