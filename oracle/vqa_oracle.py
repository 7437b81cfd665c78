"""CPU fp32 oracle for the DL_VQA training / inference step.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch (CPU, fp32), the arithmetic of the reference's
hot path.  It is imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product path (dl_vqa_b200/) never imports it.

Every function cites the reference file:line it follows (paths relative to the reference repo).

Parity pin: the reference has no golden vectors or tests, so this oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the unmodified
models/model.py, runs it on seeded inputs and commits the results under tests/golden/;
tests/test_oracle.py checks this restatement against those fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

DEFAULT_CFG = {
    # config/config.yaml:51-74 (the `train:` block; only the keys models/model.py reads)
    "text": {"question_features": 1024, "embedding_features": 300, "dropout": 0.3,
             "num_lstm_layers": 1, "bidirectional": True},
    "image": {"kernel_size": 3, "dropout": 0.3, "num_channels": [3, 64, 128, 256], "stride": 1,
              "do_skip_connection": False},
    "attention": {"hidden_dim": 1024, "glimpses": 2, "do_option": "+", "dropout": 0.3},
    "classifier": {"hidden_dim": 1024, "dropout": 0.3},
    "max_answers": 3000,
    "image_size": 224,
}


def cfg_with(base: Optional[dict] = None, **overrides) -> dict:
    """Deep-copy a config and apply 'section.key'=value overrides (e.g. **{'image.stride': 2})."""
    import copy
    cfg = copy.deepcopy(base if base is not None else DEFAULT_CFG)
    for k, val in overrides.items():
        parts = k.split(".")
        d = cfg
        for p in parts[:-1]:
            d = d[p]
        d[parts[-1]] = val
    return cfg


def zero_dropout(cfg: dict) -> dict:
    return cfg_with(cfg, **{"text.dropout": 0.0, "image.dropout": 0.0,
                            "attention.dropout": 0.0, "classifier.dropout": 0.0})


# ----------------------------------------------------------------------------------------------
# model forward, restated stage by stage
# ----------------------------------------------------------------------------------------------
def _ident(t):
    return t


def bf16_round_ste(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 storage precision with a straight-through gradient.  Used to emulate where the
    CUDA bf16 arm keeps activations in bf16 (so ReLU / max-pool gating decisions coincide)."""
    return t + (t.detach().bfloat16().float() - t.detach())


def image_encoder(sd: Dict[str, torch.Tensor], cfg: dict, v: torch.Tensor, rnd=_ident) -> torch.Tensor:
    """models/model.py:72-84 (ImageNet2): per layer Conv2d(k, stride, pad 0) -> ReLU -> MaxPool2d(2,2)
    floor mode.  The trailing Dropout is identity here (oracle = eval / dropout 0)."""
    n_layers = len(cfg["image"]["num_channels"]) - 1
    x = v
    for i in range(n_layers):
        x = F.conv2d(x, sd[f"image.conv{i}.weight"], sd[f"image.conv{i}.bias"],
                     stride=cfg["image"]["stride"])
        x = torch.clamp_min(x, 0.0)
        x = rnd(F.max_pool2d(x, kernel_size=2, stride=2))
    return x


def l2_normalise(v: torch.Tensor) -> torch.Tensor:
    """models/model.py:56: v / (||v||_2 over channels + 1e-12)."""
    n = torch.sqrt((v * v).sum(dim=1, keepdim=True))
    return v / (n + 1e-12)


def lstm_final_cell(sd, x: torch.Tensor, lengths: torch.Tensor, hidden: int,
                    bidirectional: bool, rnd=_ident, pre=None) -> torch.Tensor:
    """models/model.py:159-166: pack_padded_sequence + nn.LSTM, keep the final CELL state of every
    direction, laid out [B, dirs*H] as (c_fwd | c_bwd).  Explicit loop; gate order i,f,g,o
    (torch.nn.LSTM); forward direction consumes t=0..len-1, reverse consumes t=len-1..0.
    `pre` (test aid): per direction a [B, T, 4H] tensor of input projections W_ih x_t + b_ih + b_hh indexed by TOKEN
    position t, used instead of computing them from x -- lets a test hand the recurrence the very numbers a kernel saw
    and take gradients with respect to them."""
    B, T = x.shape[:2]
    outs = []
    for d, (suffix, reverse) in enumerate((("", False), ("_reverse", True))):
        if reverse and not bidirectional:
            break
        w_hh = sd[f"text.lstm.weight_hh_l0{suffix}"]
        if pre is None:
            w_ih = sd[f"text.lstm.weight_ih_l0{suffix}"]
            bias = sd[f"text.lstm.bias_ih_l0{suffix}"] + sd[f"text.lstm.bias_hh_l0{suffix}"]
        h = x.new_zeros(B, hidden)
        c = x.new_zeros(B, hidden)
        for s in range(T):
            active = (s < lengths)                                   # [B]
            t_idx = (lengths - 1 - s).clamp_min(0) if reverse else torch.full_like(lengths, s)
            if pre is None:
                xt = x[torch.arange(B), t_idx]                       # [B, E]
                proj = xt @ w_ih.t() + bias
            else:
                proj = pre[d][torch.arange(B), t_idx]                # [B, 4H]
            gates = rnd(proj) + h @ w_hh.t()
            gi, gf, gg, go = gates.chunk(4, dim=1)
            c_new = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h_new = rnd(torch.sigmoid(go) * torch.tanh(c_new))
            m = active.unsqueeze(1)
            c = torch.where(m, c_new, c)
            h = torch.where(m, h_new, h)
        outs.append(c)
    return torch.cat(outs, dim=1)


def lstm_live_row_blocks(row_lengths: torch.Tensor, T: int, s_begin: int = 0, block: int = 64) -> List[int]:
    """models/model.py:160-164: pack_padded_sequence hands nn.LSTM only the (step, row) positions with step < length,
    so positions past the end of a question never enter the weight gradients.  For step-indexed buffers [T][B][.]
    (row r of step s at s*B + r) this lists, in ascending order, the `block`-row blocks of steps s_begin..T-1 (indexed
    from step s_begin) that hold at least one such position -- what dl_vqa_b200's vqa_lstm_active_kblocks computes on the
    device.  With rows in descending length order step s keeps its first batch_sizes[s] rows (PackedSequence.batch_sizes),
    i.e. its first ceil(batch_sizes[s] / block) blocks."""
    B = int(row_lengths.numel())
    assert B % block == 0
    G = B // block
    lens = row_lengths.clamp(0, T).view(G, block)
    out = []
    for s in range(s_begin, T):
        for g in range(G):
            if bool((lens[g] > s).any()):
                out.append((s - s_begin) * G + g)
    return out


def question_encoder(sd, cfg: dict, q: torch.Tensor, q_len: torch.Tensor, rnd=_ident) -> torch.Tensor:
    """models/model.py:151-166: embedding (padding_idx 0) -> dropout(identity) -> tanh -> LSTM c_n."""
    emb = sd["text.embedding.weight"][q]                             # [B,T,E]
    x = rnd(torch.tanh(emb))
    return rnd(lstm_final_cell(sd, x, q_len.to(torch.long).cpu(), cfg["text"]["question_features"],
                               cfg["text"]["bidirectional"], rnd))


def attention_logits(sd, cfg: dict, v: torch.Tensor, q: torch.Tensor, rnd=_ident) -> torch.Tensor:
    """models/model.py:183-195 (Attention.forward), dropout = identity."""
    vp = rnd(F.conv2d(v, sd["attention.v_conv.weight"]))             # 1x1, no bias
    qp = q @ sd["attention.q_lin.weight"].t() + sd["attention.q_lin.bias"]
    qt = qp[:, :, None, None].expand_as(vp)                          # models/model.py:224-231
    opt = cfg["attention"]["do_option"]
    if opt == "+":
        x = torch.clamp_min(vp + qt, 0.0)
    elif opt == "*":
        x = torch.clamp_min(vp * qt, 0.0)
    elif opt == "|":
        x = torch.clamp_min(torch.cat([vp, qt], dim=1), 0.0)
    else:
        raise ValueError(f"do_option {opt!r}")
    return F.conv2d(x, sd["attention.x_conv.weight"], sd["attention.x_conv.bias"])


def glimpse_pool(v: torch.Tensor, att: torch.Tensor) -> torch.Tensor:
    """models/model.py:208-221: softmax over the spatial positions per glimpse, weighted sum of the
    (normalised) image features, flattened glimpse-major."""
    B, C = v.shape[:2]
    vf = v.reshape(B, C, -1)                                         # [B,C,S]
    p = torch.softmax(att.reshape(B, att.shape[1], -1), dim=-1)      # [B,G,S]
    return torch.einsum("bgs,bcs->bgc", p, vf).reshape(B, -1)


def classifier(sd, x: torch.Tensor, rnd=_ident) -> torch.Tensor:
    """models/model.py:198-205: drop -> lin1 -> relu -> drop -> lin2 (dropout identity)."""
    h = rnd(torch.clamp_min(x @ sd["classifier.lin1.weight"].t() + sd["classifier.lin1.bias"], 0.0))
    return h @ sd["classifier.lin2.weight"].t() + sd["classifier.lin2.bias"]


def forward(sd: Dict[str, torch.Tensor], cfg: dict, v: torch.Tensor, q: torch.Tensor,
            q_len: torch.Tensor, intermediates: Optional[dict] = None, emulate_bf16: bool = False) -> torch.Tensor:
    """models/model.py:53-67 (VqaNet.forward) in eval mode / dropout 0.  emulate_bf16 rounds the
    activations that the CUDA bf16 arm stores in bf16 (test aid; the reference itself is fp32)."""
    rnd = bf16_round_ste if emulate_bf16 else _ident
    img = image_encoder(sd, cfg, v, rnd)
    vn = rnd(l2_normalise(img))
    qf = question_encoder(sd, cfg, q, q_len, rnd)
    att = attention_logits(sd, cfg, vn, qf, rnd)
    pooled = rnd(glimpse_pool(vn, att))
    comb = torch.cat([pooled, qf], dim=1)
    logits = classifier(sd, comb, rnd)
    if intermediates is not None:
        intermediates.update(img=img, vn=vn, qf=qf, att=att, pooled=pooled, comb=comb)
    return logits


def forward_gated(sd: Dict[str, torch.Tensor], cfg: dict, v: torch.Tensor, q: torch.Tensor, q_len: torch.Tensor,
                  gates: dict) -> torch.Tensor:
    """models/model.py:53-67 with every piecewise-constant GATING decision taken from `gates` instead of from the
    activations (test aid for reduced-precision implementations).  A ReLU / max-pool network is piecewise smooth: inside
    one gating pattern the reference's gradient is a smooth function of weights and inputs, but a bf16 forward flips a
    fraction of a percent of the decisions against fp32 and every flip moves gradient entries by 100 %.  Freezing the
    pattern to the one the implementation under test took, and running the reference arithmetic in fp32 under it, gives
    the gradient that implementation must reproduce to rounding accuracy.
        gates["pool"][i]   int64 [B, C_i, PH_i, PW_i]  0..3 = (dy*2+dx) of the window element that was the maximum, 4 = ReLU-dead
        gates["att"]       bool  [B, A, S1, S2]        attention ReLU alive   (models/model.py:188-193)
        gates["cls"]       bool  [B, hidden]           classifier ReLU alive  (models/model.py:202)
    Smooth stages (L2 norm, tanh, LSTM, softmax pooling, the linear maps, the loss) are exactly forward()'s."""
    n_layers = len(cfg["image"]["num_channels"]) - 1
    x = v
    for i in range(n_layers):
        y = F.conv2d(x, sd[f"image.conv{i}.weight"], sd[f"image.conv{i}.bias"], stride=cfg["image"]["stride"])
        m = gates["pool"][i]
        B, C, PH, PW = m.shape
        win = y[:, :, :2 * PH, :2 * PW].reshape(B, C, PH, 2, PW, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, PH, PW, 4)
        x = win.gather(4, m.clamp(max=3).unsqueeze(-1)).squeeze(-1) * (m < 4).to(y.dtype)
    vn = l2_normalise(x)
    qf = question_encoder(sd, cfg, q, q_len)
    vp = F.conv2d(vn, sd["attention.v_conv.weight"])
    qp = qf @ sd["attention.q_lin.weight"].t() + sd["attention.q_lin.bias"]
    qt = qp[:, :, None, None].expand_as(vp)
    opt = cfg["attention"]["do_option"]
    if opt == "+":
        xa = (vp + qt) * gates["att"].to(vp.dtype)
    elif opt == "*":
        xa = (vp * qt) * gates["att"].to(vp.dtype)
    else:
        xa = torch.cat([vp, qt], dim=1) * gates["att"].to(vp.dtype)
    att = F.conv2d(xa, sd["attention.x_conv.weight"], sd["attention.x_conv.bias"])
    pooled = glimpse_pool(vn, att)
    comb = torch.cat([pooled, qf], dim=1)
    h = (comb @ sd["classifier.lin1.weight"].t() + sd["classifier.lin1.bias"]) * gates["cls"].to(comb.dtype)
    return h @ sd["classifier.lin2.weight"].t() + sd["classifier.lin2.bias"]


def gates_of_forward(sd, cfg, v, q, q_len) -> dict:
    """The gating pattern forward() itself takes (fp32): forward_gated(.., gates_of_forward(..)) == forward(..)."""
    n_layers = len(cfg["image"]["num_channels"]) - 1
    gates = {"pool": []}
    x = v
    for i in range(n_layers):
        y = F.conv2d(x, sd[f"image.conv{i}.weight"], sd[f"image.conv{i}.bias"], stride=cfg["image"]["stride"])
        pooled, idx = F.max_pool2d(y, 2, 2, return_indices=True)
        OW = y.shape[3]
        r, c = idx // OW, idx % OW
        e = (r % 2) * 2 + (c % 2)
        gates["pool"].append(torch.where(pooled > 0, e, torch.full_like(e, 4)))
        x = torch.clamp_min(pooled, 0.0)
    vn = l2_normalise(x)
    qf = question_encoder(sd, cfg, q, q_len)
    vp = F.conv2d(vn, sd["attention.v_conv.weight"])
    qp = qf @ sd["attention.q_lin.weight"].t() + sd["attention.q_lin.bias"]
    qt = qp[:, :, None, None].expand_as(vp)
    opt = cfg["attention"]["do_option"]
    pre = vp + qt if opt == "+" else (vp * qt if opt == "*" else torch.cat([vp, qt], dim=1))
    gates["att"] = pre > 0
    att = F.conv2d(torch.clamp_min(pre, 0.0), sd["attention.x_conv.weight"], sd["attention.x_conv.bias"])
    comb = torch.cat([glimpse_pool(vn, att), qf], dim=1)
    gates["cls"] = (comb @ sd["classifier.lin1.weight"].t() + sd["classifier.lin1.bias"]) > 0
    return gates


def step_with_grads_gated(sd, cfg, batch, gates):
    """forward_gated + loss + autograd backward.  Returns (logits, loss, grads)."""
    v, q, q_len, a_idx, a_val, _ = batch
    leaves = {k: t.detach().clone().requires_grad_(True) for k, t in sd.items()}
    logits = forward_gated(leaves, cfg, v, q, q_len, gates)
    loss = soft_target_loss_dense(logits, a_idx, a_val)
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in leaves.items()}
    grads["text.embedding.weight"][0].zero_()
    return logits.detach(), loss.detach(), grads


# ----------------------------------------------------------------------------------------------
# loss / score / optimizer (train.py, utils/train_utils.py)
# ----------------------------------------------------------------------------------------------
def soft_target_loss(logits: torch.Tensor, a_indices: torch.Tensor, a_values: torch.Tensor) -> torch.Tensor:
    """train.py:190-206.  nll = -log_softmax(y); for every non-padding answer slot (a_indices != 0)
    take nll[b, a_indices-1] * a_values/10; sum everything; divide by the batch size."""
    nll = -torch.log_softmax(logits, dim=1)
    B = logits.shape[0]
    total = logits.new_zeros(())
    for b in range(B):
        for j in range(a_indices.shape[1]):
            idx = int(a_indices[b, j])
            if idx != 0:
                total = total + nll[b, idx - 1] * (float(a_values[b, j]) / 10.0)
    return total / B


def soft_target_loss_dense(logits, a_indices, a_values) -> torch.Tensor:
    """Dense cross-check of soft_target_loss."""
    B, N = logits.shape
    tgt = torch.zeros(B, N + 1, dtype=logits.dtype)
    tgt.scatter_add_(1, a_indices.to(torch.long), a_values.to(logits.dtype) / 10.0)
    tgt = tgt[:, 1:]
    return (-torch.log_softmax(logits, dim=1) * tgt).sum() / B


def vqa_score(logits: torch.Tensor, a_indices: torch.Tensor, a_values: torch.Tensor) -> torch.Tensor:
    """utils/train_utils.py:12-25 (batch_accuracy): argmax per sample (first max), number of
    annotators that gave that answer, sum_b min(0.3*count, 1)."""
    pred = logits.argmax(dim=1)
    total = 0.0
    for b in range(logits.shape[0]):
        cnt = 0.0
        for j in range(a_indices.shape[1]):
            idx = int(a_indices[b, j])
            if idx != 0 and idx - 1 == int(pred[b]):
                cnt = float(a_values[b, j])
        total += min(0.3 * cnt, 1.0)
    return torch.tensor(total, dtype=torch.float32)


def learning_rate(initial_lr: float, iteration: int) -> float:
    """train.py:31-35."""
    return initial_lr * 0.5 ** (float(iteration) / 50000)


def adam_step(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """train.py:55,80: torch.optim.Adam defaults (no weight decay, no amsgrad).  In-place on m, v;
    returns the new parameter."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    return p - (lr / bc1) * (m / denom)


# ----------------------------------------------------------------------------------------------
# parameters and synthetic batches
# ----------------------------------------------------------------------------------------------
def param_shapes(cfg: dict, embedding_tokens: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict keys and shapes, in the reference's registration order (models/model.py:26-51)."""
    t, im, at, cl = cfg["text"], cfg["image"], cfg["attention"], cfg["classifier"]
    H, E = t["question_features"], t["embedding_features"]
    out = [("text.embedding.weight", (embedding_tokens, E))]
    for suf in ("", "_reverse") if t["bidirectional"] else ("",):
        out += [(f"text.lstm.weight_ih_l0{suf}", (4 * H, E)), (f"text.lstm.weight_hh_l0{suf}", (4 * H, H)),
                (f"text.lstm.bias_ih_l0{suf}", (4 * H,)), (f"text.lstm.bias_hh_l0{suf}", (4 * H,))]
    ch, k = im["num_channels"], im["kernel_size"]
    for i in range(len(ch) - 1):
        out += [(f"image.conv{i}.weight", (ch[i + 1], ch[i], k, k)), (f"image.conv{i}.bias", (ch[i + 1],))]
    qf = H * (2 if t["bidirectional"] else 1)
    mid, G = at["hidden_dim"], at["glimpses"]
    xin = 2 * mid if at["do_option"] == "|" else mid
    out += [("attention.v_conv.weight", (mid, ch[-1], 1, 1)), ("attention.q_lin.weight", (mid, qf)),
            ("attention.q_lin.bias", (mid,)), ("attention.x_conv.weight", (G, xin, 1, 1)),
            ("attention.x_conv.bias", (G,))]
    out += [("classifier.lin1.weight", (cl["hidden_dim"], G * ch[-1] + qf)),
            ("classifier.lin1.bias", (cl["hidden_dim"],)),
            ("classifier.lin2.weight", (cfg["max_answers"], cl["hidden_dim"])),
            ("classifier.lin2.bias", (cfg["max_answers"],))]
    return out


def random_params(cfg: dict, embedding_tokens: int, seed: int = 1, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded parameters with fan-in scaling (not the reference's init; used where the test only
    needs *some* identical weights on both sides).  Row 0 of the embedding is zero (padding_idx)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg, embedding_tokens):
        if name.endswith("bias") or ".bias_" in name:
            sd[name] = (torch.rand(shape, generator=g) - 0.5) * 0.2
        elif name == "text.embedding.weight":
            w = torch.randn(shape, generator=g)
            w[0].zero_()
            sd[name] = w
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = torch.randn(shape, generator=g) * (scale / math.sqrt(fan_in))
    return sd


def synthetic_batch(B: int, cfg: dict, embedding_tokens: int, seed: int = 1, T: int = 23, A: int = 10,
                    full_first: bool = True):
    """fp16-rounded N(0,1) images, random-length zero-padded questions with
    q_len[0] = T, sorted unique 1-based answer ids with counts summing to <= 10."""
    g = torch.Generator().manual_seed(seed)
    S = cfg.get("image_size", 224)
    v = torch.randn(B, cfg["image"]["num_channels"][0], S, S, generator=g).half().float()
    q_len = torch.randint(1, T + 1, (B,), generator=g)
    if full_first:
        q_len[0] = T
    q = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        q[b, : int(q_len[b])] = torch.randint(1, embedding_tokens, (int(q_len[b]),), generator=g)
    a_len = torch.randint(1, 5, (B,), generator=g)
    a_idx = torch.zeros(B, A, dtype=torch.long)
    a_val = torch.zeros(B, A, dtype=torch.long)
    for b in range(B):
        n = int(a_len[b])
        ids = torch.randperm(cfg["max_answers"], generator=g)[:n].sort().values + 1
        a_idx[b, :n] = ids
        cuts = torch.randint(1, 4, (n,), generator=g)
        while int(cuts.sum()) > 10:
            cuts = torch.clamp(cuts - 1, min=1)
        a_val[b, :n] = cuts
    return v, q, q_len, a_idx, a_val, a_len


def rel_err(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-12) -> float:
    """Parity metric fixed by SURVEY.md section 8d: ||a-b||_inf / (||b||_inf + eps), per tensor."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + eps))


def step_with_grads(sd, cfg, batch, requires=None, emulate_bf16: bool = False):
    """forward + loss + backward by autograd on the restated forward.  Returns
    (logits, loss, score, grads dict)."""
    v, q, q_len, a_idx, a_val, _ = batch
    leaves = {k: t.detach().clone().requires_grad_(True) for k, t in sd.items()}
    inter = {}
    logits = forward(leaves, cfg, v, q, q_len, inter, emulate_bf16=emulate_bf16)
    loss = soft_target_loss_dense(logits, a_idx, a_val)
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in leaves.items()}
    grads["text.embedding.weight"][0].zero_()       # padding_idx=0 row never receives gradient
    score = vqa_score(logits.detach(), a_idx, a_val)
    return logits.detach(), loss.detach(), score, grads, {k: t.detach() for k, t in inter.items()}
