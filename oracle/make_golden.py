"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the reference is not present on the GPU box):
    python oracle/make_golden.py

What is pinned:
  * models/model.py VqaNet forward + autograd backward (the real nn.Module, eval mode and
    dropout-0 train mode) on small configs covering do_option '+', '*', '|', stride 1/2,
    bidirectional on/off -- weights, inputs, logits, loss and every parameter gradient are stored;
  * train.py:run_batch lines 190-206 for the loss (batch_accuracy, which cannot execute on the
    installed numpy/torch -- SURVEY.md section 8c -- is stubbed out; the score stays restated);
  * the default config.yaml model at seed 1 (reference default init, V=15000): logits / loss /
    gradient digests for B=4, regenerable on any box from the seed (weights are too big to commit).
"""
import os
import sys
import types
import warnings

import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from oracle import vqa_oracle as O  # noqa: E402


def import_reference():
    # stubs for packages that are not installed here
    om = types.ModuleType("omegaconf")
    om.DictConfig = dict
    sys.modules.setdefault("omegaconf", om)
    tl = types.ModuleType("utils.train_logger")
    tl.TrainLogger = object
    sys.modules.setdefault("utils.train_logger", tl)
    from models.model import VqaNet  # noqa
    import train as ref_train  # noqa
    ref_train.batch_accuracy = lambda *a, **k: torch.tensor(0.0)
    return VqaNet, ref_train


def ref_step(VqaNet, ref_train, cfg, V, sd, batch, train_mode):
    v, q, q_len, a_idx, a_val, a_len = batch
    torch.manual_seed(0)
    m = VqaNet(cfg, V)
    if sd is not None:
        m.load_state_dict(sd)
    m.train(train_mode)
    log_softmax = torch.nn.LogSoftmax(dim=1)
    loss, _ = ref_train.run_batch(m, log_softmax, (v, q, a_idx, a_val, a_len, None, q_len), cfg["max_answers"])
    logits = m(v, q, q_len).detach()
    m.zero_grad()
    loss, _ = ref_train.run_batch(m, log_softmax, (v, q, a_idx, a_val, a_len, None, q_len), cfg["max_answers"])
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in m.named_parameters()}
    return logits, loss.detach(), grads, {k: t.detach().clone() for k, t in m.state_dict().items()}


SMALL_BASE = {
    "text": {"question_features": 16, "embedding_features": 12, "dropout": 0.0, "num_lstm_layers": 1,
             "bidirectional": True},
    "image": {"kernel_size": 3, "dropout": 0.0, "num_channels": [3, 8, 16, 32], "stride": 1,
              "do_skip_connection": False},
    "attention": {"hidden_dim": 24, "glimpses": 2, "do_option": "+", "dropout": 0.0},
    "classifier": {"hidden_dim": 20, "dropout": 0.0},
    "max_answers": 40,
    "image_size": 38,
}

SMALL_VARIANTS = {
    "plus": {},
    "mul": {"attention.do_option": "*"},
    "cat": {"attention.do_option": "|"},
    "stride2": {"image.stride": 2, "image_size": 150},
    "unidir": {"text.bidirectional": False},
    "g3": {"attention.glimpses": 3},
}


def main():
    os.makedirs(OUT, exist_ok=True)
    VqaNet, ref_train = import_reference()

    small = {}
    for name, ov in SMALL_VARIANTS.items():
        cfg = O.cfg_with(SMALL_BASE, **ov)
        V = 30
        sd = O.random_params(cfg, V, seed=7, scale=1.5)
        batch = O.synthetic_batch(2 if name == "stride2" else 5, cfg, V, seed=11, T=7, A=10)
        logits, loss, grads, _ = ref_step(VqaNet, ref_train, cfg, V, sd, batch, train_mode=True)
        logits_eval, loss_eval, _, _ = ref_step(VqaNet, ref_train, cfg, V, sd, batch, train_mode=False)
        assert torch.equal(logits, logits_eval)
        # images are fp16-representable by construction: store them as half to keep the fixture small
        assert torch.equal(batch[0].half().float(), batch[0])
        small[name] = {"cfg": cfg, "V": V, "sd": sd, "batch": (batch[0].half(),) + tuple(batch[1:]),
                       "logits": logits, "loss": loss, "grads": grads}
        print(f"small/{name}: loss {float(loss):.6f} logits {tuple(logits.shape)}")
    torch.save(small, os.path.join(OUT, "vqa_small.pt"))

    # default config at the reference's own init under seed 1
    cfg = O.zero_dropout(yaml.safe_load(open(os.path.join(REF, "config", "config.yaml")))["train"])
    V = 15000
    torch.manual_seed(1)
    m = VqaNet(cfg, V)
    sd = {k: t.detach().clone() for k, t in m.state_dict().items()}
    batch = O.synthetic_batch(4, cfg, V, seed=1)
    logits, loss, grads, _ = ref_step(VqaNet, ref_train, cfg, V, sd, batch, train_mode=True)
    digest = {k: {"absmax": float(g.abs().max()), "sum": float(g.double().sum()),
                  "l2": float(g.double().norm()), "head": g.flatten()[:16].clone()} for k, g in grads.items()}
    wdigest = {k: {"sum": float(t.double().sum()), "head": t.flatten()[:8].clone()} for k, t in sd.items()}
    full = {"cfg": cfg, "V": V, "seed": 1, "B": 4, "logits": logits, "loss": loss, "grad_digest": digest,
            "weight_digest": wdigest}
    torch.save(full, os.path.join(OUT, "vqa_full_seed1.pt"))
    print(f"full: loss {float(loss):.6f}")


if __name__ == "__main__":
    main()
