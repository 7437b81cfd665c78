"""Micro-benchmarks of BASELINE.json configs[2] and configs[3] (measurement harness, not product code).

    attention_microbench()   fused attention kernels, image-feature grid x question vector, 2 glimpses, batch 512
    lstm_microbench()        question encoder alone (embedding + tanh + bi-LSTM final cell), variable-length packed
                             questions (max 23 tokens), batch 1024 -- through VqaNet.encode_question / model.text(q, q_len)

Both time with CUDA events on the launching stream after warm-up; working sets (>= 0.9 GB / >= 0.4 GB) exceed the 126 MB L2.
bench.py embeds the results in its JSON line (`attention_microbench`, `lstm_microbench`);
`python tools/microbench.py` prints them on their own.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _time(fn, iters, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def attention_microbench(B=512, iters=20, p_drop=0.3, hbm_peak_gbs=None):
    """SURVEY.md section 8d config 3.  Algorithmic bytes per sample, forward (bf16): V' [676,1024] + q' [1024] + v [676,256]
    + out [512] = 1 733 632 B (the survey's figure; q' counted in bf16 as there -- it is fp32 here, +2 KB) => 887.6 MB per
    call.  Backward: reads V', q', v, prob, dOut; writes dV', dv, dq', per-sample dW partials."""
    import torch
    from dl_vqa_b200 import lib
    lib.load()
    P, A, C, G = 676, 1024, 256, 2
    st = lib.stream()
    g = torch.Generator(device="cuda").manual_seed(3)
    vp = torch.randn(B, P, A, device="cuda", generator=g).half()       # v' as VqaNet hands it over: fp16 from the v_conv GEMM
    qp = torch.randn(B, A, device="cuda", generator=g)
    vn = (torch.randn(B, P, C, device="cuda", generator=g) / 16).bfloat16()
    wx = torch.randn(G, A, device="cuda", generator=g) / 32
    bx = torch.zeros(G, device="cuda")
    prob = torch.empty(B, G, P, device="cuda")
    out = torch.empty(B, G * C, dtype=torch.bfloat16, device="cuda")
    dout = torch.randn(B, G * C, device="cuda", generator=g).bfloat16()
    dvp, dvn = torch.empty(B, P, A, device="cuda", dtype=torch.bfloat16), torch.empty_like(vn)
    dqp = torch.empty(B, A, device="cuda")
    dwx = torch.empty(B, G * A, device="cuda")
    dbx = torch.empty(B, G, device="cuda")
    fwd_bytes = B * (P * A + A + P * C + G * C) * 2                                            # the survey's definition
    bwd_bytes = B * ((2 * P * A + 2 * P * C + G * C) * 2 + 2 * A * 4 + G * P * 4 + G * A * 4)

    def fwd(p):
        lib.call("vqa_attention_fwd_x", lib.ptr(vp), lib.F16, lib.ptr(qp), lib.ptr(vn), lib.ptr(wx), lib.ptr(bx), lib.ptr(prob),
                 lib.ptr(out), G * C, lib.BF16, lib.ATT_ADD, B, P, A, C, G, p, 1234, st)

    def bwd(p):
        lib.call("vqa_attention_bwd_x", lib.ptr(dout), G * C, lib.ptr(vp), lib.F16, lib.ptr(qp), lib.ptr(vn), lib.ptr(wx),
                 lib.ptr(prob), lib.ptr(dvp), lib.ptr(dvn), lib.ptr(dqp), lib.ptr(dwx), lib.ptr(dbx), lib.BF16,
                 lib.ATT_ADD, B, P, A, C, G, p, 1234, st)

    res = {"config": "BASELINE.json configs[2]: fused attention, 26x26 grid x question vector, 2 glimpses, 16-bit activations "
                     "(v' fp16, v / out / gradients bf16)", "batch": B,
           "p_drop": p_drop, "fwd_bytes": fwd_bytes, "bwd_bytes": bwd_bytes,
           "l2_policy": "working set 0.9 GB (fwd) / 1.8 GB (bwd) per call exceeds the 126 MB L2"}
    for name, fn, nbytes in (("fwd_train", lambda: fwd(p_drop), fwd_bytes), ("fwd_eval", lambda: fwd(0.0), fwd_bytes),
                             ("bwd_train", lambda: bwd(p_drop), bwd_bytes), ("bwd_eval", lambda: bwd(0.0), bwd_bytes)):
        ms = _time(fn, iters)
        gbs = nbytes / ms / 1e6
        res[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1)}
        if hbm_peak_gbs:
            res[name]["frac_of_hbm_peak"] = round(gbs / hbm_peak_gbs, 4)
    return res


def lstm_microbench(B=1024, iters=10, T=23, V=15000, tflops_peak=None):
    """q [1024,23] variable length -> c_n [1024,2048]; forward, and forward + backward.
    FLOP = 2 * 2 * sum(len) * (300 + 1024) * 4096 per pass (the survey's length-aware figure; the dense T = 23 figure is
    reported beside it).  The recurrence kernels are timed on their own as well (per-step latency is what bounds them)."""
    import torch
    import dl_vqa_b200 as D
    from dl_vqa_b200 import lib, synth
    lib.load()
    cfg = synth.default_cfg(dropout=0.0)
    torch.manual_seed(1)
    m = D.VqaNet(cfg, V, compute_dtype="bfloat16").cuda().train(True)
    g = torch.Generator().manual_seed(5)
    q_len = torch.randint(1, T + 1, (B,), generator=g)
    q_len[0] = T
    q = torch.randint(1, V, (B, T), generator=g) * (torch.arange(T)[None, :] < q_len[:, None])
    q, q_len = q.cuda(), q_len.cuda()
    H, E, dirs = 1024, 300, 2
    sum_len = int(q_len.sum())
    flop_var = 2.0 * dirs * sum_len * (E + H) * 4 * H
    flop_dense = 2.0 * dirs * B * T * (E + H) * 4 * H
    dqf = torch.randn(B, dirs * H, device="cuda").bfloat16()

    def fwd():
        with torch.no_grad():
            return m.encode_question(q, q_len)

    def fwd_bwd():
        for p in m.text.parameters():
            p.grad = None
        out = m.text(q, q_len)                      # reference questionNet.forward signature
        out.backward(dqf)

    ms_f = _time(fwd, iters)
    ms_fb = _time(fwd_bwd, iters)
    tags = ["lstm_recurrence_fwd", "lstm_bwd_persistent", "lstm_inproj", "lstm_whh_wgrad", "lstm_wih_wgrad", "lstm_inproj_dgrad",
            "embed_fwd", "embed_bwd", "lstm_step_fwd", "lstm_step_bwd", "lstm_bwd_pointwise"]
    per = []
    for _ in range(5):
        lib.enable_kernel_timing(tags)
        fwd_bwd()
        per.append(lib.collect_kernel_timing())
    kern = {}
    for k in per[0]:
        xs = sorted(t[k][1] for t in per if k in t)
        kern[k] = {"calls": per[0][k][0], "ms": round(xs[len(xs) // 2], 4)}
    res = {"config": "BASELINE.json configs[3]: question encoder alone (embedding + tanh + bi-LSTM c_n), variable-length, bf16",
           "batch": B, "T": T, "mean_len": sum_len / B, "fwd_ms": round(ms_f, 4), "fwd_bwd_ms": round(ms_fb, 4),
           "fwd_questions_per_s": round(B / ms_f * 1e3, 1), "fwd_bwd_questions_per_s": round(B / ms_fb * 1e3, 1),
           "flop_per_pass_length_aware": flop_var, "flop_per_pass_dense_T": flop_dense,
           "fwd_TFLOPs_length_aware": round(flop_var / ms_f / 1e9, 1), "fwd_TFLOPs_dense": round(flop_dense / ms_f / 1e9, 1),
           "fwd_bwd_TFLOPs_dense": round(3 * flop_dense / ms_fb / 1e9, 1), "kernels": kern}
    rf = kern.get("lstm_recurrence_fwd")
    rb = kern.get("lstm_bwd_persistent")
    launches = (B + 255) // 256                      # the forward recurrence runs 256 sequences per cooperative launch
    if rf:
        res["recurrence_fwd_us_per_step"] = round(1000 * rf["ms"] / (T * launches), 2)
        res["recurrence_fwd_TFLOPs_dense"] = round(2.0 * dirs * B * T * H * 4 * H / rf["ms"] / 1e9, 1)
    if rb:
        res["recurrence_bwd_us_per_step"] = round(1000 * rb["ms"] / T, 2)
        res["recurrence_bwd_TFLOPs_dense"] = round(2.0 * dirs * B * (T - 1) * H * 4 * H / rb["ms"] / 1e9, 1)
    if tflops_peak:
        res["fwd_frac_of_tensor_peak_dense"] = round(res["fwd_TFLOPs_dense"] / tflops_peak, 4)
        res["fwd_bwd_frac_of_tensor_peak_dense"] = round(res["fwd_bwd_TFLOPs_dense"] / tflops_peak, 4)
    # SURVEY.md section 8d's secondary, "VQA-like" length distribution: clamp(round(N(6.2, 2.0)), 1, 23), one full-length
    # question kept so that the padded width stays T.  Same kernels, same shapes: what changes is how many (step, row)
    # positions the length-aware parts (tile skipping in the recurrences, block skipping in the weight gradients) leave out.
    q_len2 = torch.clamp(torch.round(torch.randn(B, generator=g) * 2.0 + 6.2), 1, T).long()
    q_len2[0] = T
    q2 = torch.randint(1, V, (B, T), generator=g) * (torch.arange(T)[None, :] < q_len2[:, None])
    q, q_len = q2.cuda(), q_len2.cuda()              # fwd / fwd_bwd close over these names
    f2, fb2 = _time(fwd, iters), _time(fwd_bwd, iters)
    res["vqa_like_lengths"] = {"distribution": "clamp(round(N(6.2, 2.0)), 1, 23), q_len[0] = 23", "mean_len": float(q_len2.float().mean()),
                               "fwd_ms": round(f2, 4), "fwd_bwd_ms": round(fb2, 4),
                               "fwd_questions_per_s": round(B / f2 * 1e3, 1), "fwd_bwd_questions_per_s": round(B / fb2 * 1e3, 1)}
    return res


if __name__ == "__main__":
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    out = {"attention_microbench": attention_microbench(hbm_peak_gbs=peaks.get("hbm_gbs")),
           "lstm_microbench": lstm_microbench(tflops_peak=peaks.get("bf16_tflops_sustained"))}
    print(json.dumps(out))
