"""Recipe for oracle/_ref/: the UNMODIFIED reference sources of the hot path, taken from where they lie under
/root/reference.  TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/).

The reference is pure Python, so "building" it is a verbatim file copy -- nothing is compiled, nothing is edited.
oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels to the
GPU box, where /root/reference does not exist.  Run by `__graft_entry__.build()` whenever /root/reference is
present; `python oracle/build_ref.py` does the same by hand.

Files (all on SURVEY.md section 8a's path):
    models/model.py            VqaNet and its sub-modules                     (the model oracle / CPU baseline)
    train.py                   run_batch, update_learning_rate, train, evaluate
    utils/train_utils.py       batch_accuracy, TrainParams, get_zeroed_metrics_dict
    utils/types.py             type aliases train.py imports
    config/config.yaml, config/config_eval.yaml      the `train:` blocks the model is built from
`oracle/ref_loader.py` imports them (with stubs for the packages that are not installed here).
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/model.py", "train.py", "utils/__init__.py", "utils/train_utils.py",
         "utils/types.py", "config/config.yaml", "config/config_eval.yaml"]


def build(verbose: bool = False) -> bool:
    """Copy the files; returns False (and leaves any existing copy alone) when /root/reference is absent."""
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("copied", rel)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref ready" if ok else f"{REF} not present; nothing copied")
    sys.exit(0 if ok or available() else 1)
