"""Micro-benchmark of vqa_tc_gemm at chosen shapes (not product code).
usage: python tools/gemm_bench.py  -> JSON {case: ms}"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_vqa_b200 import lib

lib.load()
st = lib.stream()


def run(M, N, K, ld, bias, out_dtype=torch.bfloat16, iters=10):
    A = torch.randn(M, ld, device="cuda").bfloat16()
    B = torch.randn(N, ld, device="cuda").bfloat16()
    C = torch.empty(M, N, dtype=out_dtype, device="cuda")
    b1 = torch.randn(N, device="cuda") if bias else None
    b2 = torch.randn(N, device="cuda") if bias > 1 else None
    fn = lambda: lib.call("vqa_tc_gemm", lib.ptr(A), ld, 0, lib.ptr(B), ld, 0, lib.ptr(C), lib.dtype_code(out_dtype), N, 0,
                          lib.ptr(b1), lib.ptr(b2), 0, M, N, K, 1, 0, 0.0, 0, 0, st)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / iters, 4)


only = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else None
res = {}
for name, args in {
    "inproj_K300_ld304_bias2": (5888, 4096, 300, 304, 2),
    "inproj_K300_ld304_bias0": (5888, 4096, 300, 304, 0),
    "inproj_K320_ld320_bias0": (5888, 4096, 320, 320, 0),
    "inproj_K256_ld256_bias0": (5888, 4096, 256, 256, 0),
    "inproj_K256_ld256_bias2": (5888, 4096, 256, 256, 2),
    "vconv_K256_bias0": (173056, 1024, 256, 256, 0),
    "sq_4096_K1024": (4096, 4096, 1024, 1024, 0),
    "vconv_dgrad_like_K1024_N256": (173056, 256, 1024, 1024, 0),
}.items():
    if only and name not in only:
        continue
    res[name] = run(*args)
print(json.dumps(res))
