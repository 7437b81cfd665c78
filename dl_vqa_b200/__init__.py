"""dl_vqa_b200 -- B200-native (sm_100a) training / inference step of OmerShubi/DL_VQA.

Public API (mirrors the reference):
    VqaNet(cfg, embedding_tokens)          reference models/model.py:VqaNet
    run_batch(model, log_softmax, batch, max_answers)   reference train.py:run_batch
    update_learning_rate(optimizer, iteration, initial_lr)   reference train.py:31-35
    FusedAdam(params, lr)                  reference train.py:55 (torch.optim.Adam)
    DevicePrefetcher(batches)              double-buffered pinned-host -> device copies around train.py:183-187
    GraphedTrainStep(model, optimizer, max_answers)   the loop body of train.py:69-81 as one replayed CUDA graph
"""
from .model import VqaNet  # noqa: F401
from .step import run_batch, soft_target_loss_and_score, update_learning_rate  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .pipeline import DevicePrefetcher  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401

__all__ = ["VqaNet", "run_batch", "soft_target_loss_and_score", "update_learning_rate", "FusedAdam",
           "DevicePrefetcher", "GraphedTrainStep"]
