"""One small bf16 training step for compute-sanitizer (racecheck / synccheck / memcheck) -- measurement harness, not
product code.  Shapes are chosen so that every hand-synchronised kernel of the full-size step runs: the persistent
LSTM forward and backward (H = 1024, global-counter release/acquire, mbarrier rings), both streaming attention kernels
(A = 1024, C = 256; cp.async.bulk + mbarrier ring), the tcgen05 convolutions / GEMMs (TMA + mbarrier pipelines) and
the fused Adam.  The image is 64 x 64 and the batch 4 so that a sanitizer run finishes in minutes.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dl_vqa_b200 as D  # noqa: E402
from dl_vqa_b200 import lib, synth  # noqa: E402


def main():
    lib.load()
    cfg = synth.default_cfg()
    cfg["image_size"] = int(os.environ.get("SAN_IMAGE", "64"))
    B, T, V = int(os.environ.get("SAN_BATCH", "4")), int(os.environ.get("SAN_T", "6")), 500
    torch.manual_seed(1)
    model = D.VqaNet(cfg, V, compute_dtype="bfloat16").cuda().train(True)
    opt = D.FusedAdam(model.parameters(), lr=5e-4)
    batch = tuple(t.cuda() for t in synth.make_batch(B, cfg, V, seed=2, T=T))
    n0 = lib.launch_count()
    for it in range(2):
        loss, score = D.run_batch(model, None, batch, cfg["max_answers"])
        opt.zero_grad()
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print(f"sanitize_case: 2 steps, B={B} T={T} image={cfg['image_size']}, loss {float(loss):.5f}, "
          f"{lib.launch_count() - n0} kernel launches")


if __name__ == "__main__":
    main()
