"""Micro-benchmark of the first-layer kernels (not product code): python tools/conv0_bench.py [--batch 256]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_vqa_b200 import lib
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
lib.load(); st = lib.stream()
B, IH, IW = args.batch, 224, 224
PH = PW = 111
x = torch.randn(B, 3, IH, IW, device="cuda").half().float()
w = torch.randn(64, 3, 3, 3, device="cuda") / 5; bias = torch.zeros(64, device="cuda")
out = torch.empty(B, PH, PW, 64, dtype=torch.bfloat16, device="cuda"); mask = torch.empty(B, PH, PW, 64, dtype=torch.uint8, device="cuda")
dpool = torch.randn(B, PH, PW, 64, device="cuda").bfloat16()
dw = torch.empty(64, 3, 3, 3, device="cuda"); db = torch.empty(64, device="cuda")
cases = {"fwd": (lambda: lib.call("vqa_tc_conv0_relu_pool_fwd", lib.ptr(x), lib.ptr(w), lib.ptr(bias), lib.ptr(out), lib.ptr(mask), B, IH, IW, 3, 64, st),
                 B * (3 * IH * IW * 4 + PH * PW * 64 * 3)),
         "bwd": (lambda: lib.call("vqa_tc_conv0_bwd_weight_bias", lib.ptr(x), lib.ptr(dpool), lib.ptr(mask), lib.ptr(dw), lib.ptr(db), B, IH, IW, 3, 64, st),
                 B * (3 * IH * IW * 4 + PH * PW * 64 * 3))}
res = {"batch": B}
for name, (fn, nbytes) in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.iters
    res[name] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}
print(json.dumps(res))
