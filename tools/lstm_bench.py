"""Micro-benchmark of the persistent LSTM recurrence (BASELINE.json configs[3] shape per 256-sequence launch) -- not product code."""
import argparse, os, sys, json      # VQA_LSTM_CLUSTER=4 python tools/lstm_bench.py for the cluster / multicast variant
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_vqa_b200 import lib
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
L = lib.load(); st = lib.stream()
B, T, H, dirs = args.batch, 23, 1024, 2
gx = (torch.randn(dirs, T, B, 4 * H, device="cuda") * 0.5).bfloat16()
cs = torch.empty(dirs, T, B, H, device="cuda")
hs = torch.zeros(dirs, T + 1, B, H, dtype=torch.bfloat16, device="cuda")
qf = torch.empty(B, dirs * H, dtype=torch.bfloat16, device="cuda")
w = torch.randn(dirs, 4 * H, H, device="cuda") / 32
wp = torch.empty(dirs, 4 * H, H, dtype=torch.bfloat16, device="cuda")
for d in range(dirs):
    lib.call("vqa_pack_lstm_whh", lib.ptr(w[d]), lib.ptr(wp[d]), H, st)
qlen = torch.randint(1, T + 1, (B,), device="cuda"); qlen[0] = T
sync = torch.zeros(dirs, dtype=torch.int32, device="cuda")
g0 = gx.clone()
def fn():
    gx.copy_(g0)
    lib.call("vqa_tc_lstm_fwd", lib.ptr(gx), lib.ptr(cs), lib.ptr(hs), lib.ptr(qf), lib.ptr(wp), lib.ptr(qlen), lib.ptr(sync), T, B, H, dirs, st)
for _ in range(3): fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(args.iters):
    gx.copy_(g0); a.record()
    lib.call("vqa_tc_lstm_fwd", lib.ptr(gx), lib.ptr(cs), lib.ptr(hs), lib.ptr(qf), lib.ptr(wp), lib.ptr(qlen), lib.ptr(sync), T, B, H, dirs, st)
    b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = sum(ts) / len(ts)
print(json.dumps({"batch": B, "ms": round(ms, 4), "us_per_step": round(1000 * ms / T, 2), "cluster_size": L.vqa_tc_lstm_cluster_size(),
                  "tflops": round(2.0 * dirs * T * B * H * 4 * H / ms / 1e9, 1)}))
