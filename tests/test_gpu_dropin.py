"""Drop-in proofs on the GPU (BASELINE.json north_star: "the train.py loop ... stay unchanged, so the new path is a drop-in").

* the reference's UNMODIFIED training loop -- train.py:38-141 train(), :144-169 evaluate(), :172-208 run_batch(),
  executed from oracle/_ref (verbatim copy, oracle/build_ref.py) -- is run twice on the same synthetic loaders: once
  around the reference's own VqaNet (cuDNN / cuBLAS, TF32 off) and once around dl_vqa_b200.VqaNet; the metrics it returns
  must agree.  Then once more with the fused step pieces swapped in (dl_vqa_b200.run_batch, FusedAdam through
  torch.optim.Adam's name), still inside the reference's loop.
* DevicePrefetcher hands every host batch over unchanged, in order, into a fixed set of device buffers.
"""
import types
import warnings

import pytest
import torch

from oracle import vqa_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.gpu


class _Loader(list):
    """list of batches with the two attributes train.py uses: len(loader) and len(loader.dataset)"""
    def __init__(self, batches):
        super().__init__(batches)
        self.dataset = range(sum(int(b[0].shape[0]) for b in batches))


class _Logger:
    def __init__(self):
        self.lines = []

    def __getattr__(self, name):
        def sink(*a, **k):
            self.lines.append((name, a, k))
        return sink


def _loaders(cfg, V, n_train=3, n_eval=2, B=6):
    def mk(seed):
        v, q, q_len, a_idx, a_val, a_len = O.synthetic_batch(B, cfg, V, seed=seed, T=9)
        return (v, q, a_idx, a_val, a_len, torch.arange(B), q_len)
    return _Loader([mk(100 + i) for i in range(n_train)]), _Loader([mk(200 + i) for i in range(n_eval)])


def _run_reference_loop(ref, model, cfg, loaders, epochs=2):
    params = ref.train.TrainParams(n_epochs_stop=5, num_epochs=epochs, save_model=False, max_answers=cfg["max_answers"],
                                   lr={"lr_value": 1e-3, "lr_decay": 15, "lr_gamma": 0.1, "lr_step_size": 30})
    logger = _Logger()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        metrics = ref.train.train(model, loaders[0], loaders[1], params, logger, None)
    stats = [k for (name, a, k) in logger.lines if name == "write_epoch_statistics"]
    return {k: float(v) for k, v in metrics.items()}, stats


@pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built (python oracle/build_ref.py)")
def test_unmodified_reference_training_loop_around_the_drop_in_model():
    import dl_vqa_b200 as D
    ref = ref_loader.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.zero_dropout(O.cfg_with(O.DEFAULT_CFG, image_size=64))
    V = 300
    sd = O.random_params(cfg, V, seed=4, scale=1.5)
    loaders = _loaders(cfg, V)

    m_ref = ref.VqaNet(cfg, V)
    m_ref.load_state_dict(sd)
    m_ref.cuda().train(True)
    want, want_stats = _run_reference_loop(ref, m_ref, cfg, loaders)

    m = D.VqaNet(cfg, V)                                   # exact (fp32) arm: the reference's arithmetic
    m.load_state_dict(sd)
    m.cuda().train(True)
    got, got_stats = _run_reference_loop(ref, m, cfg, loaders)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 2e-3 * max(1.0, abs(want[k])), (k, got, want)
    assert len(got_stats) == len(want_stats) == 2
    for a, b in zip(got_stats, want_stats):
        assert abs(float(a["train_loss"]) - float(b["train_loss"])) <= 2e-3 * abs(float(b["train_loss"]))
        assert abs(float(a["eval_score"]) - float(b["eval_score"])) < 1e-3
    assert m.training                                      # train() leaves the model in train mode (train.py:105)
    # the checkpoint the reference would write loads into the reference's own module
    ref.VqaNet(cfg, V).load_state_dict(m.state_dict())

    # same loop with the fused pieces swapped in by name: fused loss / score, FusedAdam
    m2 = D.VqaNet(cfg, V, compute_dtype="bfloat16")
    m2.load_state_dict(sd)
    m2.cuda().train(True)
    class _TorchWithFusedAdam:                              # `torch.optim.Adam` as train.py:55 spells it -> FusedAdam
        optim = types.SimpleNamespace(Adam=D.FusedAdam)

        def __getattr__(self, k):
            return getattr(torch, k)

    real_run_batch, real_torch = ref.train.run_batch, ref.train.torch
    try:
        ref.train.run_batch = D.run_batch
        ref.train.torch = _TorchWithFusedAdam()
        got2, _ = _run_reference_loop(ref, m2, cfg, loaders)
    finally:
        ref.train.run_batch = real_run_batch
        ref.train.torch = real_torch
    for k in want:
        assert abs(got2[k] - want[k]) <= 5e-2 * max(1.0, abs(want[k])), (k, got2, want)


def test_device_prefetcher_hands_batches_over_unchanged():
    import dl_vqa_b200 as D
    cfg = O.cfg_with(O.DEFAULT_CFG, image_size=32)
    host = []
    for i in range(5):
        v, q, q_len, a_idx, a_val, a_len = O.synthetic_batch(4, cfg, 50, seed=i, T=6)
        host.append(tuple(t.pin_memory() for t in (v, q, a_idx, a_val, a_len)) + (None, q_len.pin_memory()))
    pf = D.DevicePrefetcher(host)
    ptrs = set()
    for epoch in range(2):                                  # re-iterating reuses the same two device buffer sets
        n = 0
        for hb, db in zip(host, pf):
            for h, d in zip(hb, db):
                if h is None:
                    assert d is None
                else:
                    assert d.is_cuda and d.dtype == h.dtype and torch.equal(d.cpu(), h)
            ptrs.add(db[0].data_ptr())
            n += 1
        assert n == len(host)
    assert len(ptrs) == 2
    assert pf.h2d_bytes == 2 * sum(t.numel() * t.element_size() for hb in host for t in hb if t is not None)
    # float16 hand-over: the loader yields float32, the prefetcher narrows (exactly) and copies half the bytes
    pf16 = D.DevicePrefetcher(host, image_dtype=torch.float16)
    for hb, db in zip(host, pf16):
        assert db[0].dtype == torch.float16 and torch.equal(db[0].float().cpu(), hb[0])
        assert torch.equal(db[1].cpu(), hb[1])
    img_bytes = sum(hb[0].numel() for hb in host)
    assert pf16.h2d_bytes == pf.h2d_bytes // 2 - 2 * img_bytes
    bad = [(torch.full((2, 3, 8, 8), 1.0 + 2.0 ** -15),) + host[0][1:]]
    with pytest.raises(ValueError):
        for _ in D.DevicePrefetcher(bad, image_dtype=torch.float16):
            pass
