"""Micro-benchmark of the fused attention kernels (BASELINE.json configs[2]: batch 512, 2 glimpses) -- not product code.
usage: python tools/att_bench.py [--batch 512] [--iters 10] [--p 0.3]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_vqa_b200 import lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--p", type=float, default=0.3)
ap.add_argument("--only", default="")
args = ap.parse_args()
lib.load()
B, P, A, C, G = args.batch, 676, 1024, 256, 2
st = lib.stream()
vp = torch.randn(B, P, A, device="cuda").bfloat16()
qp = torch.randn(B, A, device="cuda")
vn = torch.randn(B, P, C, device="cuda").bfloat16()
wx = torch.randn(G, A, device="cuda") / 32
bx = torch.zeros(G, device="cuda")
prob = torch.empty(B, G, P, device="cuda")
out = torch.empty(B, G * C, dtype=torch.bfloat16, device="cuda")
dout = torch.randn(B, G * C, device="cuda").bfloat16()
dvp, dvn = torch.empty_like(vp), torch.empty_like(vn)
dqp = torch.empty(B, A, device="cuda"); dwx = torch.empty(B, G * A, device="cuda"); dbx = torch.empty(B, G, device="cuda")
fwd_bytes = B * ((P * A + P * C + G * C) * 2 + A * 4 + G * P * 4)
bwd_bytes = B * ((2 * P * A + 2 * P * C + G * C) * 2 + 2 * A * 4 + G * P * 4 + G * A * 4)
cases = {
    "fwd": (lambda: lib.call("vqa_attention_fwd", lib.ptr(vp), lib.ptr(qp), lib.ptr(vn), lib.ptr(wx), lib.ptr(bx), lib.ptr(prob),
                             lib.ptr(out), G * C, lib.BF16, lib.ATT_ADD, B, P, A, C, G, args.p, 1234, st), fwd_bytes),
    "bwd": (lambda: lib.call("vqa_attention_bwd", lib.ptr(dout), G * C, lib.ptr(vp), lib.ptr(qp), lib.ptr(vn), lib.ptr(wx),
                             lib.ptr(prob), lib.ptr(dvp), lib.ptr(dvn), lib.ptr(dqp), lib.ptr(dwx), lib.ptr(dbx), lib.BF16,
                             lib.ATT_ADD, B, P, A, C, G, args.p, 1234, st), bwd_bytes),
}
res = {"batch": B, "p_drop": args.p}
for name, (fn, nbytes) in cases.items():
    if args.only and name not in args.only.split(","):
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
    res[name] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "bytes": nbytes}
print(json.dumps(res))
