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
        self._shadow = {}          # id(param) -> (bf16 tensor, validity entry [tensor, param, version]) : VqaNet.use_weight_shadows
        self._dev_state = None     # device-resident VqaStepState (lr / bias corrections read on the GPU): use_device_step_state()

    _RING = 4

    def register_bf16_shadow(self, param, shadow, entry):
        """`shadow` (bf16, same element order as `param`) receives the updated parameter in every step();
        `entry[2]` is set to the parameter's version afterwards so that the model can tell a current shadow."""
        if shadow.numel() != param.numel() or shadow.dtype != torch.bfloat16 or not shadow.is_contiguous():
            raise ValueError("bf16 shadow must be a contiguous bf16 tensor with the parameter's element count")
        self._shadow[id(param)] = (shadow, entry)

    def use_device_step_state(self, state):
        """Read lr and the Adam step count from a device-resident `VqaStepState` (include/vqa_b200.h) that `vqa_step_tick`
        advances, instead of from `param_groups[..]['lr']` and the host-side step counter: the optimizer step can then be
        part of a captured CUDA graph (dl_vqa_b200/graph.py).  The host counters keep being advanced as well, so
        `state_dict()` stays meaningful.  None switches back."""
        self._dev_state = state
        return self

    def after_replay(self):
        """Host bookkeeping for one optimizer step that ran inside a replayed CUDA graph (no kernel is launched): advance the
        per-parameter step counters, bump the parameters' autograd versions and re-validate the bf16 shadows."""
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] += 1
                    torch.autograd.graph.increment_version(p)
                    sh = self._shadow.get(id(p))
                    if sh is not None:
                        sh[1][2] = p._version

    def _table(self, gi, tensors_key, lists):
        """Device-resident pointer tables, rebuilt only when a pointer changes (gradient tensors usually do change
        from step to step).  The upload goes through a small ring of PINNED host buffers with non-blocking copies:
        a pageable `.to(device)` would block the host until the stream drains, i.e. synchronise host and GPU once
        per step and expose the launch latency of every following forward pass."""
        cached = self._tables.get(gi)
        if cached is not None and cached["key"] == tensors_key:
            return cached["dev"][cached["slot"]]
        dev = lists["p"][0].device
        n = len(lists["p"])
        if cached is None or cached["n"] != n:
            cached = {"n": n, "slot": -1, "key": None,
                      "host": [torch.empty(6, n, dtype=torch.int64).pin_memory() for _ in range(self._RING)],
                      "dev": [torch.empty(6, n, dtype=torch.int64, device=dev) for _ in range(self._RING)],
                      "done": [None] * self._RING}
            self._tables[gi] = cached
        slot = (cached["slot"] + 1) % self._RING
        if cached["done"][slot] is not None:
            cached["done"][slot].synchronize()          # the copy that last used this pinned buffer (4 steps ago)
        host = cached["host"][slot]
        host.copy_(torch.tensor([[t.data_ptr() for t in lists[k]] for k in ("p", "g", "m", "v")] +
                                [[t.numel() for t in lists["p"]], lists["s"]], dtype=torch.int64))
        cached["dev"][slot].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        cached["done"][slot] = ev
        cached["slot"], cached["key"] = slot, tensors_key
        return cached["dev"][slot]

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            lists = {"p": [], "g": [], "m": [], "v": [], "s": []}
            shadowed = []
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
                sh = self._shadow.get(id(p))
                lists["s"].append(sh[0].data_ptr() if sh is not None else 0)
                if sh is not None:
                    shadowed.append((p, sh[1]))
            step = self.state[ps[0]]["step"]
            key = tuple(t.data_ptr() for k in ("p", "g", "m", "v") for t in lists[k]) + tuple(lists["s"])
            table = self._table(gi, key, lists)
            n = len(ps)
            b1, b2 = group["betas"]
            if self._dev_state is not None:
                call("vqa_adam_multi_dev", ptr(table[0]), ptr(table[1]), ptr(table[2]), ptr(table[3]),
                     ptr(table[5]) if shadowed else None, ptr(table[4]),
                     n, max(t.numel() for t in lists["p"]), ptr(self._dev_state), float(b1), float(b2), float(group["eps"]),
                     float(grad_scale), lib.stream())
            else:
                call("vqa_adam_multi", ptr(table[0]), ptr(table[1]), ptr(table[2]), ptr(table[3]),
                     ptr(table[5]) if shadowed else None, ptr(table[4]),
                     n, max(t.numel() for t in lists["p"]), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                     int(step), float(grad_scale), lib.stream())
            for p in ps:                       # the kernel wrote through p.data: tell autograd (VqaNet's backward checks it)
                torch.autograd.graph.increment_version(p)
            for p, entry in shadowed:          # the kernel has just rewritten these shadows from the updated masters
                entry[2] = p._version
        return loss
