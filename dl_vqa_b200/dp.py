"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce of bucketed gradients over
NVLink/NVSwitch, overlapped with the rest of backward.

The reference is single-GPU (main.py:23); this layer is new.  VqaNet's backward finishes parameter groups
in a fixed order (classifier 22.8 MB -> attention 9.4 MB -> question encoder 61 MB -> image encoder 1.5 MB)
and fires `grad_ready_hook` after each group; every group becomes one bucket that is all-reduced
asynchronously on NCCL's stream while the remaining (conv-dominated) backward runs.  `finish()` makes the
compute stream wait for the reductions.

Two paths, chosen per backward pass:
  * in place -- with the model's gradient arena (`VqaNet.use_gradient_arena`, switched on by this wrapper when
    available) a stage's gradients ARE one flat buffer: the all-reduce runs on it directly, there is no
    concatenate / scatter-back copy;
  * copying  -- when the backward handed out fresh tensors (no arena, or a live `.grad` still aliased the arena:
    gradient accumulation, a missed `zero_grad`, `set_to_none=False`), the bucket is a concatenated copy; after
    the reduction `finish()` corrects `p.grad` itself (looked up by NAME after backward: autograd's AccumulateGrad
    may have cloned the tensor the hook saw, or added it into an existing `.grad`) by `reduced - local`.
The SCALE convention is fixed at construction and never changes from step to step: with an arena the gradients hold
SUMS over the ranks and the division by the world size is folded into the optimizer
(`FusedAdam.step(grad_scale=ddp.grad_scale)`, grad_scale = 1/world) -- on both paths, so accumulating a copied
bucket onto an in-place one stays consistent; without an arena `finish()` divides and grad_scale is 1.
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
        self.sum_convention = False      # gradients hold sums, the optimizer divides (arena available)
        self.in_place = False            # path taken by the most recent backward pass (informational)
        self.wait_events = None          # optional [(before, after)] CUDA events around finish()'s stream waits
        if self.world > 1:
            model.grad_ready_hook = self._on_group_ready
            if hasattr(model, "use_gradient_arena"):
                model.use_gradient_arena(True)
                self.sum_convention = True
                self.in_place = True

    @property
    def grad_scale(self) -> float:
        """Factor the optimizer must apply to the gradients: 1/world when they hold sums (arena available and
        averaging requested), else 1 (finish() has already divided, or sums were asked for)."""
        return 1.0 / self.world if (self.sum_convention and self.average and self.world > 1) else 1.0

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
        flat = self._bucket_of(named_grads) if self.sum_convention else None
        if flat is not None:                                          # all-reduce the stage's bucket in place
            self.in_place = True
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._pending.append((work, None, None, None))
            return
        # fresh tensors this step: reduce a concatenated copy.  Only NAMES are kept -- holding the tensors would make
        # AccumulateGrad clone them, and the tensor that ends up in p.grad is looked up after backward anyway.
        self.in_place = False
        local = torch.cat([g.detach().reshape(-1) for _, g in named_grads])
        red = local.clone()
        work = dist.all_reduce(red, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self._pending.append((work, red, local, [(n, g.numel()) for n, g in named_grads]))

    def finish(self):
        """Call after loss.backward(): waits (stream-wise) for the buckets; on the copying path also replaces the local
        contribution inside every `p.grad` by the reduced one."""
        if not self._pending:
            return
        ev = None
        if self.wait_events is not None and torch.cuda.is_available():
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        params = None
        for work, red, local, names in self._pending:
            work.wait()
            if red is None:
                continue                                              # in place: nothing to scatter, scale in the optimizer
            if self.average and not self.sum_convention:
                red.div_(self.world)
            if params is None:
                params = dict(self.model.named_parameters())
            red.sub_(local)                                           # p.grad = (whatever it accumulated) - local + reduced
            off = 0
            for n, cnt in names:
                p = params.get(n)
                if p is not None and p.grad is not None:
                    p.grad.add_(red[off:off + cnt].view_as(p.grad))
                off += cnt
        self._pending.clear()
        if ev is not None:
            ev[1].record()
            self.wait_events.append(ev)

    def broadcast_parameters(self, src: int = 0):
        if self.world > 1:
            with torch.no_grad():
                for p in self.model.parameters():
                    dist.broadcast(p.data, src=src, group=self.pg)
                    p.add_(0)        # bump the autograd version: bf16 weight shadows keyed on it are now stale
