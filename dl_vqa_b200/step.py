"""Loss / score tail of the training step (reference train.py:172-208) on the fused CUDA kernels.

`run_batch(model, log_softmax, batch_data, max_answers)` keeps the reference's signature and return
values `(batch_loss, batch_score)`; the loss participates in autograd.  Unlike the reference it performs
no device->host synchronisation (the reference builds gather indices with numpy and calls .item() per
sample, train.py:194-199 and utils/train_utils.py:21-23).
"""
from __future__ import annotations

import torch

from . import lib
from .lib import call, ptr


class _SoftTargetLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, a_indices, a_values):
        B, N = logits.shape
        A = a_indices.shape[1]
        dev = logits.device
        logits = logits.contiguous()
        dlogits = torch.empty_like(logits)
        rows = torch.empty(2, B, dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        call("vqa_softloss_fwd_bwd", ptr(logits), ptr(a_indices), ptr(a_values), ptr(dlogits), ptr(rows[0]),
             ptr(rows[1]), ptr(out[0:1]), ptr(out[1:2]), B, N, A, lib.stream())
        ctx.save_for_backward(dlogits)
        loss, score = out[0], out[1]                  # bind the views ONCE: the object marked is the object returned
        ctx.mark_non_differentiable(score)
        return loss, score

    @staticmethod
    def backward(ctx, g_loss, g_score):
        (dlogits,) = ctx.saved_tensors
        # d loss / d logits was produced by the forward kernel for g_loss = 1; scale by the incoming gradient, which is a
        # device scalar (no host read, no ATen kernel)
        g = g_loss if (g_loss.dtype == torch.float32 and g_loss.is_contiguous()) else g_loss.to(torch.float32).contiguous()
        out = torch.empty_like(dlogits)
        call("vqa_scale_by_device_scalar", ptr(dlogits), ptr(out), ptr(g), dlogits.numel(), lib.stream())
        return out, None, None


def soft_target_loss_and_score(logits: torch.Tensor, a_indices: torch.Tensor, a_values: torch.Tensor):
    """train.py:190-206 loss and utils/train_utils.py:12-25 score from one fused kernel.
    logits fp32 [B,N] on CUDA; a_indices/a_values int64 [B,A] (1-based ids, 0 = padding)."""
    if not logits.is_cuda:
        raise lib.VqaLibraryError("soft_target_loss_and_score: CUDA tensors only (no CPU fallback)")
    dev = logits.device
    a_indices = a_indices.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
    a_values = a_values.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
    return _SoftTargetLoss.apply(logits.to(torch.float32), a_indices, a_values)


def run_batch(model, log_softmax, batch_data, max_answers):
    """Drop-in for reference train.py:172-208.  `log_softmax` is accepted and ignored (fused)."""
    v, q, a_indices, a_values, a_length, idx, q_len = batch_data
    dev = next(model.parameters()).device
    v = v.to(dev, non_blocking=True)
    q = q.to(dev, non_blocking=True)
    q_len = torch.as_tensor(q_len).to(dev, non_blocking=True)
    y_hat = model(v, q, q_len)
    if y_hat.shape[1] != max_answers:
        raise ValueError(f"model produces {y_hat.shape[1]} answers, max_answers={max_answers}")
    loss, score = soft_target_loss_and_score(y_hat, a_indices, a_values)
    if not torch.is_grad_enabled():
        # evaluation (train.py:144-169 runs under @torch.no_grad()): the reference's batch_accuracy returns a CPU tensor
        # and evaluate() accumulates it into `score = torch.tensor(0.0)` (a CPU tensor, train.py:155,165) -- hand the
        # score over on the host there.  In training the loop adds it to a Python number (train.py:87), so the device
        # tensor is kept and the step stays free of host synchronisation.
        score = score.cpu()
    return loss, score


def update_learning_rate(optimizer, iteration, initial_lr):
    """reference train.py:31-35: lr = lr0 * 0.5 ** (iteration / 50000)."""
    lr = initial_lr * 0.5 ** (float(iteration) / 50000)
    for group in optimizer.param_groups:
        group["lr"] = lr
    return lr
