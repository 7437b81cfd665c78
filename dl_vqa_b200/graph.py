"""The whole training step as ONE CUDA graph.

The reference's loop body (train.py:69-81)

    batch_loss, batch_score = run_batch(model, log_softmax, batch_data, max_answers)
    optimizer.zero_grad()
    update_learning_rate(optimizer, total_iterations, lr)
    batch_loss.backward()
    optimizer.step()
    total_iterations += 1

is ~75 kernel launches here; enqueuing them one by one costs the host 2-3 ms per 4.5 ms step, which leaves no slack
once eight ranks share a host.  `GraphedTrainStep` captures that body once per distinct set of input buffers --
forward, fused loss/score, backward, the data-parallel all-reduce of every gradient bucket (NCCL is captured with the
rest), multi-tensor Adam -- and replays it with a single launch.  Everything the host used to compute per step lives
in a 64-byte device-resident `VqaStepState` (include/vqa_b200.h) that the first node of the graph (`vqa_step_tick`)
advances: the dropout seed, the iteration count, the learning rate of train.py:31-35 and Adam's bias corrections.

    step = GraphedTrainStep(model, optimizer, max_answers, ddp=ddp, lr=5e-4)
    for batch in DevicePrefetcher(loader):          # device tensors in a small, fixed set of buffers
        loss, score = step(batch)                   # device scalars (overwritten by the next call on the same buffers)

Every call performs exactly one training step: the first call on a new set of buffers runs eagerly (that also
initialises NCCL / lazy library state), the second one captures (capturing executes nothing) and replays.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import lib
from .lib import call, ptr
from .step import run_batch


class GraphedTrainStep:
    def __init__(self, model, optimizer, max_answers: int, ddp=None, lr: float = 5e-4, half_life: float = 50000.0,
                 iteration: int = 0, enabled: bool = True):
        if not hasattr(optimizer, "use_device_step_state"):
            raise TypeError("GraphedTrainStep needs a dl_vqa_b200.FusedAdam")
        self.model, self.opt, self.ddp = model, optimizer, ddp
        self.max_answers = int(max_answers)
        self.lr0, self.half_life = float(lr), float(half_life)
        self.enabled = enabled
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise lib.VqaLibraryError("GraphedTrainStep: the model must be on a CUDA device (no CPU fallback)")
        self.device = dev
        self.iteration = int(iteration)
        adam_step = max([int(optimizer.state[p]["step"]) for g in optimizer.param_groups for p in g["params"]
                         if optimizer.state.get(p)] or [0])
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) & ((1 << 62) - 1)
        host = torch.zeros(8, dtype=torch.int64)
        host[0], host[1], host[2] = seed, self.iteration, adam_step
        self.state = host.to(dev)                                   # VqaStepState (64 bytes)
        model.use_gradient_arena(True)                              # stable gradient addresses: required for replay
        model.use_device_step_state(self.state)
        optimizer.use_device_step_state(self.state)
        self._seen: Dict[tuple, int] = {}
        self._graphs: Dict[tuple, dict] = {}
        self._pool = None
        self.launches_per_replay = 0
        self.replays = 0

    # ------------------------------------------------------------------ one step, enqueued kernel by kernel
    def _eager(self, batch):
        b1, b2 = self.opt.param_groups[0]["betas"]
        call("vqa_step_tick", ptr(self.state), self.lr0, self.half_life, float(b1), float(b2), lib.stream())
        loss, score = run_batch(self.model, None, batch, self.max_answers)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        scale = 1.0
        if self.ddp is not None:
            self.ddp.finish()
            scale = self.ddp.grad_scale
        self.opt.step(grad_scale=scale)
        return loss.detach(), score.detach()

    def _host_lr(self):
        lr = self.lr0 * 0.5 ** (float(self.iteration) / self.half_life)           # mirror of the device value, for observers
        for g in self.opt.param_groups:
            g["lr"] = lr

    @staticmethod
    def _key(batch: Sequence) -> tuple:
        return tuple((t.data_ptr(), tuple(t.shape), t.dtype) if torch.is_tensor(t) else None for t in batch)

    def _capture(self, batch, key):
        steps = {p: self.opt.state[p]["step"] for g in self.opt.param_groups for p in g["params"] if self.opt.state.get(p)}
        n0 = lib.launch_count()
        g = torch.cuda.CUDAGraph()
        kw = {"pool": self._pool} if self._pool is not None else {}
        torch.cuda.synchronize(self.device)
        with torch.cuda.graph(g, capture_error_mode="thread_local", **kw):
            loss, score = self._eager(batch)
        if self._pool is None:
            self._pool = g.pool()
        self.launches_per_replay = lib.launch_count() - n0
        for p, s in steps.items():                      # capture ran the host code of optimizer.step() but no kernel
            self.opt.state[p]["step"] = s
        entry = {"graph": g, "loss": loss, "score": score, "batch": tuple(batch)}
        self._graphs[key] = entry
        return entry

    # ------------------------------------------------------------------ public
    def __call__(self, batch):
        """batch = (v, q, a_indices, a_values, a_length, index, q_len) of DEVICE tensors (train.py:182)."""
        for t in batch:
            if torch.is_tensor(t) and not t.is_cuda:
                raise lib.VqaLibraryError("GraphedTrainStep: device tensors only (wrap the loader in DevicePrefetcher)")
        self._host_lr()
        key = self._key(batch)
        entry = self._graphs.get(key) if self.enabled else None
        if entry is None and self.enabled and self._seen.get(key, 0) >= 1:
            entry = self._capture(batch, key)
        if entry is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            out = self._eager(batch)
        else:
            entry["graph"].replay()
            self.opt.after_replay()
            self.replays += 1
            out = (entry["loss"], entry["score"])
        self.iteration += 1
        return out

    def eager_step(self, batch):
        """One training step enqueued kernel by kernel (never captured): profiling passes, debugging."""
        self._host_lr()
        out = self._eager(batch)
        self.iteration += 1
        return out

    def read_state(self) -> dict:
        """Device state copied to the host (synchronises): for tests and checkpoints."""
        h = self.state.cpu()
        f = h[3:5].view(torch.float32)            # bytes 24..39: lr, lr_over_bc1, inv_sqrt_bc2, reserved
        return {"seed": int(h[0]), "iteration": int(h[1]), "adam_step": int(h[2]), "lr": float(f[0]),
                "lr_over_bc1": float(f[1]), "inv_sqrt_bc2": float(f[2])}
