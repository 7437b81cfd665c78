"""Micro-benchmark of the tcgen05 3x3 conv kernels at the config.yaml layer shapes (not product code).
usage: python tools/conv_bench.py [--batch 256] [--iters 10] [--cta-group 1|2] [--only fwd1,dg1,...]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_vqa_b200 import lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cta-group", type=int, default=1)
ap.add_argument("--only", default="")
args = ap.parse_args()
lib.load()
lib.call("vqa_tc_conv_set_cta_group", args.cta_group)
B = args.batch
st = lib.stream()
layers = {1: (111, 111, 64, 128), 2: (54, 54, 128, 256)}
only = set(args.only.split(",")) if args.only else None
res = {}
for li, (IH, IW, Cin, Cout) in layers.items():
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    x = torch.randn(B, IH, IW, Cin, device="cuda").bfloat16()
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)
    bias = torch.zeros(Cout, device="cuda")
    wp = torch.empty(Cout, 9 * Cin, dtype=torch.bfloat16, device="cuda")
    wd = torch.empty(Cin, 9 * Cout, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_pack_conv3x3_weight", lib.ptr(w), lib.ptr(wp), lib.ptr(wd), Cout, Cin, st)
    out = torch.empty(B, PH, PW, Cout, dtype=torch.bfloat16, device="cuda")
    mask = torch.empty(B, PH, PW, Cout, dtype=torch.uint8, device="cuda")
    dy = torch.randn(B, 2 * PH, 2 * PW, Cout, device="cuda").bfloat16()
    dx = torch.empty(B, IH, IW, Cin, dtype=torch.bfloat16, device="cuda")
    dw = torch.empty(Cout, Cin, 3, 3, device="cuda")
    flop = 2.0 * B * (IH - 2) * (IW - 2) * Cout * Cin * 9
    mask_b = torch.randint(0, 5, (B, IH, IW, Cin), device="cuda", dtype=torch.uint8)     # pooling mask of the layer below
    dy_b = torch.empty(B, 2 * IH, 2 * IW, Cin, dtype=torch.bfloat16, device="cuda")
    db_b = torch.empty(Cin, device="cuda")
    cases = {
        f"dgu{li}": lambda: lib.call("vqa_tc_conv3x3_bwd_data_unpool", lib.ptr(dy), lib.ptr(wd), lib.ptr(mask_b), lib.ptr(dy_b),
                                     lib.ptr(db_b), B, IH, IW, Cin, Cout, st),
        f"up{li}": lambda: lib.call("vqa_unpool_bf16", lib.ptr(dx), lib.ptr(mask_b), lib.ptr(dy_b), lib.ptr(db_b),
                                    B, IH, IW, Cin, st),
        f"fwd{li}": lambda: lib.call("vqa_tc_conv3x3_relu_pool_fwd", lib.ptr(x), lib.ptr(wp), lib.ptr(bias), lib.ptr(out),
                                     lib.ptr(mask), B, IH, IW, Cin, Cout, st),
        f"dg{li}": lambda: lib.call("vqa_tc_conv3x3_bwd_data", lib.ptr(dy), lib.ptr(wd), lib.ptr(dx), B, IH, IW, Cin, Cout, st),
        f"wg{li}": lambda: lib.call("vqa_tc_conv3x3_bwd_weight", lib.ptr(x), lib.ptr(dy), lib.ptr(dw), B, IH, IW, Cin, Cout, st),
    }
    for name, fn in cases.items():
        if only and name not in only:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        res[name] = {"ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1)}
if not only or "dnu" in only:      # fused L2-norm/dropout backward + un-pool of the last layer vs the two separate kernels
    PH = PW = 26; C = 256; R = B * PH * PW
    v = torch.randn(R, C, device="cuda").bfloat16()
    vn, vnd, nrm = torch.empty_like(v), torch.empty_like(v), torch.empty(R, device="cuda")
    lib.call("vqa_dropnorm_fwd", lib.ptr(v), lib.ptr(vn), lib.ptr(vnd), lib.ptr(nrm), lib.BF16, R, C, 0.1, 0.2, 7, st)
    g1, g2 = torch.randn(R, C, device="cuda").bfloat16(), torch.randn(R, C, device="cuda").bfloat16()
    mk = torch.randint(0, 5, (R, C), device="cuda", dtype=torch.uint8)
    da = torch.empty_like(v)
    dyl = torch.empty(B, 2 * PH, 2 * PW, C, dtype=torch.bfloat16, device="cuda")
    dbl = torch.empty(C, device="cuda")
    def two():
        lib.call("vqa_dropnorm_bwd", lib.ptr(g1), lib.ptr(g2), lib.ptr(vn), lib.ptr(nrm), lib.ptr(da), lib.BF16, R, C, 0.1, 0.2, 7, st)
        lib.call("vqa_unpool_bf16", lib.ptr(da), lib.ptr(mk), lib.ptr(dyl), lib.ptr(dbl), B, PH, PW, C, st)
    def one():
        lib.call("vqa_dropnorm_bwd_unpool", lib.ptr(g1), lib.ptr(g2), lib.ptr(vn), lib.ptr(nrm), lib.ptr(mk), lib.ptr(dyl),
                 lib.ptr(dbl), B, PH, PW, C, 0.1, 0.2, 7, st)
    for name, fn in (("dn+up", two), ("dnu", one)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = {"ms": round(a.elapsed_time(b) / args.iters, 4)}
print(json.dumps({"cta_group": args.cta_group, "batch": B, **res}))
