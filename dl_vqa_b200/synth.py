"""Synthetic VQA-2.0-shaped batches: there is no dataset offline.

Shapes and dtypes follow what reference preprocessing/data_preprocessing.py:74-87 yields per sample:
fp16-rounded normalised image [3,S,S] widened to fp32, zero-padded int64 question [T] with its length,
sorted unique 1-based answer ids [A] (0 padded) with their annotator counts.
"""
from __future__ import annotations

import torch

DEFAULT_CFG = {
    # reference config/config.yaml:51-74
    "text": {"question_features": 1024, "embedding_features": 300, "dropout": 0.3, "num_lstm_layers": 1,
             "bidirectional": True},
    "image": {"kernel_size": 3, "dropout": 0.3, "num_channels": [3, 64, 128, 256], "stride": 1,
              "do_skip_connection": False},
    "attention": {"hidden_dim": 1024, "glimpses": 2, "do_option": "+", "dropout": 0.3},
    "classifier": {"hidden_dim": 1024, "dropout": 0.3},
    "max_answers": 3000,
    "image_size": 224,
}
DEFAULT_TOKENS = 15000
DEFAULT_T = 23


def default_cfg(dropout=None):
    import copy
    cfg = copy.deepcopy(DEFAULT_CFG)
    if dropout is not None:
        for k in ("text", "image", "attention", "classifier"):
            cfg[k]["dropout"] = float(dropout)
    return cfg


def make_batch(B, cfg=None, embedding_tokens=DEFAULT_TOKENS, seed=1, T=DEFAULT_T, A=10, pin=False):
    """Returns host tensors (v, q, a_indices, a_values, a_length, index, q_len) in the order of the
    reference Dataset.__getitem__ / train.py:182."""
    cfg = cfg or DEFAULT_CFG
    g = torch.Generator().manual_seed(seed)
    S = cfg.get("image_size", 224)
    v = torch.randn(B, cfg["image"]["num_channels"][0], S, S, generator=g).half().float()
    q_len = torch.randint(1, T + 1, (B,), generator=g)
    q_len[0] = T
    q = torch.randint(1, embedding_tokens, (B, T), generator=g)
    q = q * (torch.arange(T)[None, :] < q_len[:, None])
    a_len = torch.randint(1, 5, (B,), generator=g)
    ids = torch.rand(B, cfg["max_answers"], generator=g).argsort(dim=1)[:, :A] + 1
    ids = ids * (torch.arange(A)[None, :] < a_len[:, None])
    ids = torch.where(ids > 0, ids, torch.full_like(ids, 1 << 40)).sort(dim=1).values
    ids = torch.where(ids >= (1 << 40), torch.zeros_like(ids), ids)
    vals = torch.randint(1, 3, (B, A), generator=g) * (ids > 0)
    out = (v, q, ids, vals, a_len, torch.arange(B), q_len)
    if pin:
        out = tuple(t.pin_memory() for t in out)
    return out
