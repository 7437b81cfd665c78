"""Host-side (Python) cost of one training step: cProfile over a few steps (not product code)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dl_vqa_b200 as D
from dl_vqa_b200 import lib, synth
lib.load()
dev = torch.device("cuda", 0)
cfg = synth.default_cfg()
torch.manual_seed(1)
model = D.VqaNet(cfg, synth.DEFAULT_TOKENS, compute_dtype="bfloat16").to(dev).train(True)
opt = D.FusedAdam(model.parameters(), lr=5e-4)
model.use_gradient_arena(True); model.use_weight_shadows(opt)
hv, hq, hai, hav, hal, _, hql = synth.make_batch(256, cfg, seed=1, pin=True)
db = tuple(t.to(dev) for t in (hv, hq, hai, hav, hal, hql))
def step():
    dv, dq, dai, dav, dal, dql = db
    loss, score = D.run_batch(model, None, (dv, dq, dai, dav, dal, None, dql), cfg["max_answers"])
    opt.zero_grad(set_to_none=True)
    D.update_learning_rate(opt, 0, 5e-4)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1000*(t1-t0)/5:.2f} ms/step; with final sync {1000*(t2-t0)/5:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
