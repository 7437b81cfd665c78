"""Fused multi-tensor Adam (reference train.py:55,76-80 uses torch.optim.Adam with defaults).

One kernel launch updates every parameter tensor (vqa_adam_multi).  Same constructor arguments and
state_dict layout idea as torch.optim.Adam (state: step, exp_avg, exp_avg_sq), no weight decay / amsgrad.
"""
from __future__ import annotations

import torch

from . import lib
from .lib import call, ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps)
        super().__init__(params, defaults)
        self._tables = {}

    def _table(self, gi, tensors_key, lists):
        """Device-resident pointer tables, rebuilt only when a pointer changes."""
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == tensors_key:
            return cached[1]
        dev = lists["p"][0].device
        host = torch.tensor([[t.data_ptr() for t in lists[k]] for k in ("p", "g", "m", "v")] +
                            [[t.numel() for t in lists["p"]]], dtype=torch.int64)
        table = host.to(dev)
        self._tables[gi] = (tensors_key, table)
        return table

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            lists = {"p": [], "g": [], "m": [], "v": []}
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise lib.VqaLibraryError("FusedAdam: fp32 CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                lists["p"].append(p.data); lists["g"].append(g); lists["m"].append(st["exp_avg"])
                lists["v"].append(st["exp_avg_sq"])
            step = self.state[ps[0]]["step"]
            key = tuple(t.data_ptr() for k in ("p", "g", "m", "v") for t in lists[k])
            table = self._table(gi, key, lists)
            n = len(ps)
            b1, b2 = group["betas"]
            call("vqa_adam_multi", ptr(table[0]), ptr(table[1]), ptr(table[2]), ptr(table[3]), None, ptr(table[4]),
                 n, max(t.numel() for t in lists["p"]), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                 int(step), float(grad_scale), lib.stream())
        return loss
