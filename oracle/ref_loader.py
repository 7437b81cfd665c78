"""Import the unmodified reference (oracle/_ref/, see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY.

    ref = load()            # namespace: VqaNet, train (the reference's train.py module), train_cfg(name)
    ref.VqaNet(cfg, V)      # reference models/model.py:VqaNet
    ref.train.run_batch / ref.train.train / ref.train.evaluate / ref.train.update_learning_rate

Packages the reference imports but this image lacks are stubbed: `omegaconf` (only used for a
type annotation) and `utils.train_logger` (TensorBoard).  ONE function cannot execute on the installed numpy / torch
and is replaced: `utils/train_utils.py:12-25 batch_accuracy` (its numpy-array indexing of a tensor raises, SURVEY.md
section 8a row a12); the replacement is the restated `oracle.vqa_oracle.vqa_score`, returning a CPU tensor as the
original does.  Everything else -- the model, run_batch's loss (train.py:190-206), the LR schedule, the epoch loop and
evaluate() -- is the reference's own code, executed as is.
"""
import importlib
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
_cached = None


def available() -> bool:
    return os.path.exists(os.path.join(REFDIR, "models", "model.py")) and os.path.exists(os.path.join(REFDIR, "train.py"))


def load():
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    import torch
    from oracle import vqa_oracle as O
    om = types.ModuleType("omegaconf")
    om.DictConfig = dict
    sys.modules.setdefault("omegaconf", om)
    # the reference's top-level package names are generic (`utils`, `models`, `train`): import them under a private
    # path entry and take them out of sys.modules again so that nothing else in the process can pick them up
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("utils", "models", "train") or
             k.startswith(("utils.", "models."))}
    tl = types.ModuleType("utils.train_logger")
    tl.TrainLogger = object
    sys.path.insert(0, REFDIR)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            importlib.invalidate_caches()
            import utils as ref_utils                      # noqa: F401  (oracle/_ref/utils)
            sys.modules["utils.train_logger"] = tl
            model_mod = importlib.import_module("models.model")
            train_mod = importlib.import_module("train")
    finally:
        sys.path.remove(REFDIR)
        for k in [k for k in sys.modules if k in ("utils", "models", "train") or k.startswith(("utils.", "models."))]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    assert os.path.abspath(model_mod.__file__).startswith(REFDIR), model_mod.__file__

    def batch_accuracy(predicted, true):
        indices, values, _size = true
        return O.vqa_score(predicted.detach().float().cpu(), indices.cpu(), values.cpu())

    train_mod.batch_accuracy = batch_accuracy

    def train_cfg(name="config.yaml"):
        import yaml
        return yaml.safe_load(open(os.path.join(REFDIR, "config", name)))["train"]

    _cached = types.SimpleNamespace(VqaNet=model_mod.VqaNet, model=model_mod, train=train_mod, train_cfg=train_cfg,
                                    torch=torch)
    return _cached
