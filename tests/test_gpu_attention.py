"""Fused attention kernels (reference models/model.py:187-195 + :208-221) through the C ABI:
generic kernels against a plain torch fp32 restatement (eval), the streaming bf16 kernels against the generic
kernels on identical inputs with the SAME dropout seed (both regenerate the same counter-based mask)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_forward(vp, qp, vn, wx, bx, op):
    """vp [B,P,A], qp [B,A], vn [B,P,C], wx [G,A], bx [G] -> prob [B,G,P], out [B,G*C]  (fp32 torch)"""
    x = vp + qp[:, None, :] if op == "+" else vp * qp[:, None, :]
    x = torch.relu(x)
    logit = torch.einsum("bpa,ga->bgp", x, wx) + bx[None, :, None]
    prob = torch.softmax(logit, dim=2)
    out = torch.einsum("bgp,bpc->bgc", prob, vn).flatten(1)
    return prob, out


def _call_fwd(vp, qp, vn, wx, bx, op, p_drop=0.0, seed=0):
    from dl_vqa_b200 import lib
    B, P, A = vp.shape
    C, G = vn.shape[2], wx.shape[0]
    dt = lib.dtype_code(vp.dtype)
    prob = torch.empty(B, G, P, device="cuda")
    out = torch.empty(B, G * C, dtype=vp.dtype, device="cuda")
    lib.call("vqa_attention_fwd", lib.ptr(vp), lib.ptr(qp), lib.ptr(vn), lib.ptr(wx), lib.ptr(bx), lib.ptr(prob),
             lib.ptr(out), G * C, dt, lib.ATT_ADD if op == "+" else lib.ATT_MUL, B, P, A, C, G, p_drop, seed, lib.stream())
    torch.cuda.synchronize()
    return prob, out


def _call_bwd(dout, vp, qp, vn, wx, prob, op, p_drop=0.0, seed=0):
    from dl_vqa_b200 import lib
    B, P, A = vp.shape
    C, G = vn.shape[2], wx.shape[0]
    dt = lib.dtype_code(vp.dtype)
    dvp, dvn = torch.empty_like(vp), torch.empty_like(vn)
    dqp = torch.empty(B, A, device="cuda")
    dwx = torch.empty(B, G * A, device="cuda")
    dbx = torch.empty(B, G, device="cuda")
    lib.call("vqa_attention_bwd", lib.ptr(dout), G * C, lib.ptr(vp), lib.ptr(qp), lib.ptr(vn), lib.ptr(wx), lib.ptr(prob),
             lib.ptr(dvp), lib.ptr(dvn), lib.ptr(dqp), lib.ptr(dwx), lib.ptr(dbx), dt,
             lib.ATT_ADD if op == "+" else lib.ATT_MUL, B, P, A, C, G, p_drop, seed, lib.stream())
    torch.cuda.synchronize()
    return dvp, dvn, dqp, dwx.sum(0).view(G, A), dbx.sum(0)


def _inputs(B, P, A, C, G, seed, dtype):
    g = torch.Generator(device="cuda").manual_seed(seed)
    vp = torch.randn(B, P, A, device="cuda", generator=g).to(dtype)
    qp = torch.randn(B, A, device="cuda", generator=g) * 0.7
    vn = (torch.randn(B, P, C, device="cuda", generator=g) / C ** 0.5).to(dtype)
    wx = torch.randn(G, A, device="cuda", generator=g) / A ** 0.5
    bx = torch.randn(G, device="cuda", generator=g)
    return vp, qp, vn, wx, bx


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20))


@pytest.mark.parametrize("op", ["+", "*"])
@pytest.mark.parametrize("B,P,A,C,G", [(3, 676, 1024, 256, 2), (2, 37, 64, 32, 3), (5, 9, 1024, 256, 1)])
def test_generic_fp32_forward_backward_match_torch(B, P, A, C, G, op):
    vp, qp, vn, wx, bx = _inputs(B, P, A, C, G, 11 + P, torch.float32)
    prob, out = _call_fwd(vp, qp, vn, wx, bx, op)
    vp_, qp_, vn_, wx_, bx_ = (t.clone().requires_grad_(True) for t in (vp, qp, vn, wx, bx))
    rprob, rout = _ref_forward(vp_, qp_, vn_, wx_, bx_, op)
    assert _rel(prob, rprob) < 1e-5 and _rel(out, rout) < 1e-5
    dout = torch.randn_like(out)
    rout.backward(dout)
    dvp, dvn, dqp, dwx, dbx = _call_bwd(dout, vp, qp, vn, wx, prob, op)
    assert _rel(dvp, vp_.grad) < 1e-4 and _rel(dvn, vn_.grad) < 1e-4 and _rel(dqp, qp_.grad) < 1e-4
    assert _rel(dwx, wx_.grad) < 1e-4 and float((dbx - bx_.grad).abs().max()) < 1e-5


@pytest.mark.parametrize("op", ["+", "*"])
@pytest.mark.parametrize("B,P,G,p_drop", [(300, 676, 2, 0.3), (7, 676, 2, 0.0), (3, 41, 1, 0.3), (149, 8, 2, 0.5), (2, 1000, 2, 0.3)])
def test_streaming_bf16_forward_matches_generic(B, P, G, p_drop, op):
    """The tensor-core arm's kernel (bf16, A=1024, C=256) against the generic kernel run in fp32 on the same
    bf16-rounded inputs, same dropout seed: both must drop exactly the same elements."""
    A, C = 1024, 256
    vp, qp, vn, wx, bx = _inputs(B, P, A, C, G, 5 + B, torch.bfloat16)
    seed = 0xC0FFEE + B
    prob, out = _call_fwd(vp, qp, vn, wx, bx, op, p_drop, seed)
    rprob, rout = _call_fwd(vp.float(), qp, vn.float(), wx, bx, op, p_drop, seed)
    # q' is rounded to bf16 and the fusion runs in bf16 in the streaming kernel: 2^-9 relative per element
    assert float((prob - rprob).abs().max()) < 2e-2 * float(rprob.max()), float((prob - rprob).abs().max())
    assert _rel(out, rout) < 2e-2, _rel(out, rout)
    if p_drop > 0:
        # a different seed must change the result; eval (p = 0) must differ from train
        _, out2 = _call_fwd(vp, qp, vn, wx, bx, op, p_drop, seed + 1)
        assert not torch.equal(out, out2)


@pytest.mark.parametrize("op", ["+", "*"])
@pytest.mark.parametrize("B,P,G,p_drop", [(300, 676, 2, 0.3), (5, 676, 2, 0.0), (3, 41, 1, 0.3), (149, 8, 2, 0.5), (2, 700, 2, 0.3)])
def test_streaming_bf16_backward_matches_generic(B, P, G, p_drop, op):
    """Streaming backward (bf16) against the generic fp32 kernel on the same bf16-rounded inputs and dropout seed.
    q' is pre-rounded to bf16 so that both kernels gate exactly the same elements (the streaming kernels add in bf16)."""
    A, C = 1024, 256
    vp, qp, vn, wx, bx = _inputs(B, P, A, C, G, 9 + B, torch.bfloat16)
    qp = qp.bfloat16().float()
    seed = 0xBEEF + B
    prob, _ = _call_fwd(vp.float(), qp, vn.float(), wx, bx, op, p_drop, seed)
    dout = (torch.randn(B, G * C, device="cuda")).bfloat16()
    got = _call_bwd(dout, vp, qp, vn, wx, prob, op, p_drop, seed)
    want = _call_bwd(dout.float(), vp.float(), qp, vn.float(), wx, prob, op, p_drop, seed)
    names = ["dvp", "dvn", "dqp", "dwx", "dbx"]
    for n, g_, w_ in zip(names, got, want):
        err = _rel(g_, w_)
        assert err < 2e-2, (n, err)
    # the gate must be identical: dvp is exactly zero in the same places (ReLU-dead or dropped)
    assert bool(((got[0].float() == 0) == (want[0] == 0)).float().mean() > 0.999)


# ------------------------------------------------------------------------------------------------------------------
# train mode (dropout on) against a torch restatement that consumes the SAME mask, read back through vqa_dropout_mask
# ------------------------------------------------------------------------------------------------------------------
def _keep_mask(B, P, A, p_drop, seed):
    from dl_vqa_b200 import lib
    keep = torch.empty(B * P * A, dtype=torch.uint8, device="cuda")
    SITE_ATT_X = 4
    lib.call("vqa_dropout_mask", lib.ptr(keep), keep.numel(), p_drop, seed, SITE_ATT_X, 1, lib.stream())
    return keep.view(B, P, A).bool()


def _ste_bf16(t):
    return t + (t.detach().bfloat16().float() - t.detach())


def _ref_forward_train(vp, qp, vn, wx, bx, op, keep, p_drop, bf16_fusion):
    """models/model.py:187-195 + 208-221 with the dropout of :194 given as an explicit keep mask.  bf16_fusion restates the
    streaming kernel's arithmetic: q' rounded to bf16, fusion result rounded to bf16 (straight-through gradients)."""
    q = _ste_bf16(qp) if bf16_fusion else qp
    x = vp + q[:, None, :] if op == "+" else vp * q[:, None, :]
    if bf16_fusion:
        x = _ste_bf16(x)
    x = torch.relu(x) * keep.float() / (1.0 - p_drop)
    logit = torch.einsum("bpa,ga->bgp", x, wx) + bx[None, :, None]
    prob = torch.softmax(logit, dim=2)
    out = torch.einsum("bgp,bpc->bgc", prob, vn).flatten(1)
    return prob, out


@pytest.mark.parametrize("op", ["+", "*"])
@pytest.mark.parametrize("dtype,B,P,A,C,G,p_drop", [
    (torch.bfloat16, 300, 676, 1024, 256, 2, 0.3),      # streaming kernels, more samples than CTAs
    (torch.bfloat16, 3, 41, 1024, 256, 1, 0.4),
    (torch.bfloat16, 7, 676, 1024, 256, 2, 0.0),
    (torch.float32, 3, 676, 1024, 256, 2, 0.3),         # generic kernels (exact arm)
    (torch.float32, 2, 37, 64, 32, 3, 0.5),
])
def test_train_mode_forward_and_backward_match_torch_with_the_same_mask(dtype, B, P, A, C, G, p_drop, op):
    vp, qp, vn, wx, bx = _inputs(B, P, A, C, G, 31 + B + P, dtype)
    seed = 0xD0D0 + B
    streaming = dtype == torch.bfloat16
    keep = _keep_mask(B, P, A, p_drop, seed) if p_drop > 0 else torch.ones(B, P, A, dtype=torch.bool, device="cuda")
    if p_drop > 0 and keep.numel() > 1000000:
        assert abs(float(keep.float().mean()) - (1 - p_drop)) < 5e-3
    prob, out = _call_fwd(vp, qp, vn, wx, bx, op, p_drop, seed)
    leaves = [t.detach().float().clone().requires_grad_(True) for t in (vp, qp, vn, wx, bx)]
    rprob, rout = _ref_forward_train(*leaves, op, keep, p_drop, bf16_fusion=streaming)
    tol = 2e-2 if streaming else 1e-4
    assert float((prob - rprob).abs().max()) < tol * float(rprob.max())
    assert _rel(out, rout) < tol, _rel(out, rout)
    dout = torch.randn(B, G * C, device="cuda").to(dtype)
    rout.backward(dout.float())
    dvp, dvn, dqp, dwx, dbx = _call_bwd(dout, vp, qp, vn, wx, prob, op, p_drop, seed)
    for name, got, want in (("dvp", dvp, leaves[0].grad), ("dvn", dvn, leaves[2].grad), ("dqp", dqp, leaves[1].grad),
                            ("dwx", dwx, leaves[3].grad)):
        assert _rel(got, want) < (2e-2 if streaming else 2e-4), (name, _rel(got, want))
    assert float((dbx - leaves[4].grad).abs().max()) < (1e-2 if streaming else 1e-4) * max(1.0, float(dwx.abs().max()))
    # the kernels gate exactly the elements the restatement gates: dropped or ReLU-dead <=> zero gradient
    dead = leaves[0].grad == 0
    assert float(((dvp.float() == 0) == dead).float().mean()) > 0.9995


# ------------------------------------------------------------------------------------------------------------------
# float16 v' (what VqaNet hands the streaming kernels: the v_conv GEMM writes fp16, model.py) -- vqa_attention_*_x
# ------------------------------------------------------------------------------------------------------------------
def _call_x(vp, qp, vn, wx, bx, op, p_drop, seed, dout=None, prob=None):
    from dl_vqa_b200 import lib
    B, P, A = vp.shape
    C, G = vn.shape[2], wx.shape[0]
    vdt = lib.F16 if vp.dtype == torch.float16 else lib.BF16
    opc = lib.ATT_ADD if op == "+" else lib.ATT_MUL
    if dout is None:
        prob = torch.empty(B, G, P, device="cuda")
        out = torch.empty(B, G * C, dtype=torch.bfloat16, device="cuda")
        lib.call("vqa_attention_fwd_x", lib.ptr(vp), vdt, lib.ptr(qp), lib.ptr(vn), lib.ptr(wx), lib.ptr(bx), lib.ptr(prob),
                 lib.ptr(out), G * C, lib.BF16, opc, B, P, A, C, G, p_drop, seed, lib.stream())
        torch.cuda.synchronize()
        return prob, out
    dvp = torch.empty(B, P, A, dtype=torch.bfloat16, device="cuda")
    dvn = torch.empty_like(vn)
    dqp = torch.empty(B, A, device="cuda")
    dwx = torch.empty(B, G * A, device="cuda")
    dbx = torch.empty(B, G, device="cuda")
    lib.call("vqa_attention_bwd_x", lib.ptr(dout), G * C, lib.ptr(vp), vdt, lib.ptr(qp), lib.ptr(vn), lib.ptr(wx), lib.ptr(prob),
             lib.ptr(dvp), lib.ptr(dvn), lib.ptr(dqp), lib.ptr(dwx), lib.ptr(dbx), lib.BF16, opc, B, P, A, C, G, p_drop, seed,
             lib.stream())
    torch.cuda.synchronize()
    return dvp, dvn, dqp, dwx.sum(0).view(G, A), dbx.sum(0)


@pytest.mark.parametrize("op", ["+", "*"])
@pytest.mark.parametrize("B,P,G,p_drop", [(300, 676, 2, 0.3), (3, 41, 1, 0.4), (7, 676, 2, 0.0)])
def test_streaming_kernels_with_float16_vprime_match_torch_and_beat_the_bf16_form(B, P, G, p_drop, op):
    """v' in fp16, q' ten times larger than the spatial variation of v' (as at the reference's initialisation): against the
    EXACT fp32 fusion on the same inputs and mask (no rounding restated) the fp16 form must hold 5e-3 on every output --
    the bf16 form of the same kernels is only held to 2e-2 against a restatement that rounds like it does."""
    from dl_vqa_b200 import lib
    A, C = 1024, 256
    assert lib.load().vqa_attention_streaming_ok(lib.BF16, lib.ATT_ADD, P, A, C, G) == 1
    g = torch.Generator(device="cuda").manual_seed(77 + B)
    base = torch.randn(B, 1, A, device="cuda", generator=g)                       # per-channel level of v', constant over positions
    vp32 = base + 0.3 * torch.randn(B, P, A, device="cuda", generator=g)
    qp = 3.0 * torch.randn(B, A, device="cuda", generator=g)
    vn = (torch.randn(B, P, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    wx = torch.randn(G, A, device="cuda", generator=g) / A ** 0.5
    bx = torch.randn(G, device="cuda", generator=g)
    seed = 0xFACE + B
    keep = _keep_mask(B, P, A, p_drop, seed) if p_drop > 0 else torch.ones(B, P, A, dtype=torch.bool, device="cuda")
    dout = torch.randn(B, G * C, device="cuda").bfloat16()
    errs = {}
    for name, vp in (("f16", vp32.half()), ("bf16", vp32.bfloat16())):
        prob, out = _call_x(vp, qp, vn, wx, bx, op, p_drop, seed)
        leaves = [t.detach().float().clone().requires_grad_(True) for t in (vp, qp, vn, wx, bx)]
        rprob, rout = _ref_forward_train(*leaves, op, keep, p_drop, bf16_fusion=False)      # exact fusion arithmetic
        rout.backward(dout.float())
        dvp, dvn, dqp, dwx, dbx = _call_x(vp, qp, vn, wx, bx, op, p_drop, seed, dout=dout, prob=prob)
        errs[name] = {"prob": float((prob - rprob).abs().max() / rprob.max()), "out": _rel(out, rout), "dwx": _rel(dwx, leaves[3].grad),
                      "dqp": _rel(dqp, leaves[1].grad), "dvn": _rel(dvn, leaves[2].grad)}
    for k, e in errs["f16"].items():
        if k == "dqp":
            # dq' = sum_s dpre[s] of gradients the kernel has rounded to bf16 for the dv' output; with a per-channel level of
            # v' that is constant over the positions (this test's construction) the exact sum nearly cancels, so a
            # relative bound says nothing here -- dq' is held to 2e-2 on unstructured inputs by the tests above
            continue
        assert e < (5e-3 if k not in ("out", "dvn") else 8e-3), (k, errs)       # out / dvn are themselves stored in bf16
    if op == "+":
        assert errs["f16"]["dwx"] < 0.5 * errs["bf16"]["dwx"], errs             # the point of the fp16 hand-over
