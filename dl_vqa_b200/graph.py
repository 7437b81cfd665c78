"""The whole training step as replayed CUDA graphs.

The reference's loop body (train.py:69-81)

    batch_loss, batch_score = run_batch(model, log_softmax, batch_data, max_answers)
    optimizer.zero_grad()
    update_learning_rate(optimizer, total_iterations, lr)
    batch_loss.backward()
    optimizer.step()
    total_iterations += 1

is ~65 kernel launches here; enqueuing them one by one costs the host 2-3 ms per 4.5 ms step, which leaves no slack
once eight ranks share a host.  `GraphedTrainStep` captures that body once per distinct set of input buffers and
replays it: forward, fused loss / score, backward and multi-tensor Adam run straight on the C ABI (no autograd engine,
no per-step Python arithmetic).  Everything the host used to compute per step lives in a 64-byte device-resident
`VqaStepState` (include/vqa_b200.h) that the first node of the graph (`vqa_step_tick`) advances: the dropout seed, the
iteration count, the learning rate of train.py:31-35 and Adam's bias corrections.

Data parallel (world > 1): NCCL stays OUTSIDE the graphs.  The capture is cut where `VqaNet._run_backward` reports that
the classifier / attention / text gradients are complete, so one step is

    graph A (forward, loss, backward down to the question encoder)
    all-reduce of the arena slice [classifier | attention | text]  (93.5 MB, one NCCL call, asynchronous)
    graph B (convolution backward: 40 % of the step, overlaps the all-reduce)
    all-reduce of the image slice (1.5 MB), stream waits
    graph C (Adam, gradients scaled by 1 / world)

-- three graph launches and two NCCL calls per step from the host.

    step = GraphedTrainStep(model, optimizer, max_answers, ddp=ddp, lr=5e-4)
    for batch in DevicePrefetcher(loader):          # device tensors in a small, fixed set of buffers
        loss, score = step(batch)                   # device scalars (overwritten by the next call on the same buffers)

Every call performs exactly one training step: the first call on a new set of buffers runs kernel by kernel (that also
initialises lazy library state), the second one captures (capturing executes nothing) and replays.
"""
from __future__ import annotations

import gc
from typing import Dict, Sequence

import torch
import torch.distributed as dist

from . import lib
from .lib import call, ptr


class GraphedTrainStep:
    def __init__(self, model, optimizer, max_answers: int, ddp=None, lr: float = 5e-4, half_life: float = 50000.0,
                 iteration: int = 0, enabled: bool = True):
        if not hasattr(optimizer, "use_device_step_state"):
            raise TypeError("GraphedTrainStep needs a dl_vqa_b200.FusedAdam")
        self.model, self.opt = model, optimizer
        self.max_answers = int(max_answers)
        self.lr0, self.half_life = float(lr), float(half_life)
        self.enabled = enabled
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise lib.VqaLibraryError("GraphedTrainStep: the model must be on a CUDA device (no CPU fallback)")
        self.device = dev
        self.iteration = int(iteration)
        self.world = ddp.world if ddp is not None else 1
        self.pg = ddp.pg if ddp is not None else None
        self.grad_scale = 1.0 / self.world if (ddp is None or ddp.average) else 1.0
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
        self._pool = torch.cuda.graph_pool_handle()                 # one private memory pool shared by every captured segment
        self._stream = torch.cuda.Stream(device=dev)                # capture stream
        self.launches_per_replay = 0
        self.replays = 0
        self.wait_events = None          # optional list: (before, after) CUDA events around the all-reduce waits (world > 1)

    # ------------------------------------------------------------------ the step on the C ABI, in program order
    def _program(self, batch, cut):
        """Enqueue one step.  `cut(kind)` is called at the two points where data-parallel communication belongs:
        cut("head") when the classifier / attention / text gradients are complete, cut("image") when all are."""
        model, opt = self.model, self.opt
        v, q, a_idx, a_val, _a_len, _idx, q_len = batch
        st = lib.stream()
        b1, b2 = opt.param_groups[0]["betas"]
        call("vqa_step_tick", ptr(self.state), self.lr0, self.half_life, float(b1), float(b2), st)
        dev = self.device
        q_len = torch.as_tensor(q_len).to(device=dev, dtype=torch.int64).contiguous()
        q = q.to(device=dev, dtype=torch.int64).contiguous()
        v = model._image_input(v)
        seed = lib.SEED_ON_DEVICE | self.state.data_ptr()
        hook, model.grad_ready_hook = model.grad_ready_hook, None            # communication is placed by `cut`, not by hooks
        try:
            logits, ctx = model._run_forward(v, q, q_len, seed, save=True)
            B, N = logits.shape
            if N != self.max_answers:
                raise ValueError(f"model produces {N} answers, max_answers={self.max_answers}")
            a_idx = a_idx.to(device=dev, dtype=torch.int64).contiguous()
            a_val = a_val.to(device=dev, dtype=torch.int64).contiguous()
            dlogits = torch.empty_like(logits)
            rows = torch.empty(2, B, dtype=torch.float32, device=dev)
            out = torch.empty(2, dtype=torch.float32, device=dev)
            call("vqa_softloss_fwd_bwd", ptr(logits), ptr(a_idx), ptr(a_val), ptr(dlogits), ptr(rows[0]), ptr(rows[1]),
                 ptr(out[0:1]), ptr(out[1:2]), B, N, a_idx.shape[1], st)
            for p in model.parameters():                         # optimizer.zero_grad(set_to_none=True)
                p.grad = None
            grads = model._run_backward(ctx, dlogits, before_image=lambda: cut("head"))
        finally:
            model.grad_ready_hook = hook
        for n, p in model.named_parameters():
            p.grad = grads[n]
        cut("image")
        opt.step(grad_scale=self.grad_scale)
        return out[0], out[1]

    # ------------------------------------------------------------------ communication (world > 1), always outside graphs
    def _slices(self):
        a = self.model._arena
        b = a["buckets"]
        head_n = b["classifier"].numel() + b["attention"].numel() + b["text"].numel()
        return {"head": a["whole"][:head_n], "image": b["image"]}

    def _comm(self, kind, pending):
        if self.world <= 1:
            return
        flat = self._slices()[kind]
        pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        if kind == "image":                                       # everything is in flight: the optimizer needs all of it
            ev = None
            if self.wait_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            for w in pending:
                w.wait()
            pending.clear()
            if ev is not None:
                ev[1].record()
                self.wait_events.append(ev)

    def _eager(self, batch):
        pending = []
        return self._program(batch, lambda kind: self._comm(kind, pending))

    # ------------------------------------------------------------------ capture: one graph per segment between cuts
    def _capture(self, batch, key):
        steps = {p: self.opt.state[p]["step"] for g in self.opt.param_groups for p in g["params"] if self.opt.state.get(p)}
        n0 = lib.launch_count()
        segs = []
        torch.cuda.synchronize(self.device)
        gc.collect()
        cur = torch.cuda.current_stream(self.device)
        self._stream.wait_stream(cur)

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=self._pool, capture_error_mode="thread_local")
            segs.append(g)

        def cut(kind):
            if self.world <= 1:
                return                                            # single GPU: the whole step is one graph
            segs[-1].capture_end()
            segs.append(kind)
            begin()

        with torch.cuda.stream(self._stream):
            begin()
            try:
                loss, score = self._program(batch, cut)
            finally:
                segs[-1].capture_end()
        cur.wait_stream(self._stream)
        self.launches_per_replay = lib.launch_count() - n0
        for p, s in steps.items():                      # capture ran the host code of optimizer.step() but no kernel
            self.opt.state[p]["step"] = s
        entry = {"segments": segs, "loss": loss, "score": score, "batch": tuple(batch)}
        self._graphs[key] = entry
        return entry

    def _replay(self, entry):
        pending = []
        for seg in entry["segments"]:
            if isinstance(seg, str):
                self._comm(seg, pending)
            else:
                seg.replay()
        self.opt.after_replay()
        self.replays += 1
        return entry["loss"], entry["score"]

    def _host_lr(self):
        lr = self.lr0 * 0.5 ** (float(self.iteration) / self.half_life)           # mirror of the device value, for observers
        for g in self.opt.param_groups:
            g["lr"] = lr

    @staticmethod
    def _key(batch: Sequence) -> tuple:
        return tuple((t.data_ptr(), tuple(t.shape), t.dtype) if torch.is_tensor(t) else None for t in batch)

    # ------------------------------------------------------------------ public
    def __call__(self, batch):
        """batch = (v, q, a_indices, a_values, a_length, index, q_len) of DEVICE tensors (train.py:182).
        Returns (loss, score) as device scalars, like run_batch."""
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
            out = self._replay(entry)
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
