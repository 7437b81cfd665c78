"""N>1 host logic on CPU: world_size-2 gloo run of the gradient bucketing / averaging layer (dl_vqa_b200/dp.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(5, 3))
        self.b = torch.nn.Parameter(torch.zeros(7))
        self.grad_ready_hook = None


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dl_vqa_b200.dp import GradientAllReduce
    m = _FakeModel()
    with torch.no_grad():
        m.a.fill_(float(rank + 1))
    ddp = GradientAllReduce(m)
    assert m.grad_ready_hook is not None
    ddp.broadcast_parameters()
    assert float(m.a[0, 0]) == 1.0                       # rank 0's value everywhere
    # two "stages" finishing at different times, as VqaNet._run_backward fires them
    g1 = [("b", torch.full((7,), float(rank))), ("a", torch.full((5, 3), 10.0 * (rank + 1)))]
    g2 = [("c", torch.arange(4, dtype=torch.float32) * (rank + 1))]
    ddp._on_group_ready(g1)
    ddp._on_group_ready(g2)
    ddp.finish()
    ok = (torch.allclose(g1[0][1], torch.full((7,), 0.5)) and torch.allclose(g1[1][1], torch.full((5, 3), 15.0))
          and torch.allclose(g2[0][1], torch.arange(4, dtype=torch.float32) * 1.5) and not ddp._pending)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_single_process_is_a_noop():
    from dl_vqa_b200.dp import GradientAllReduce
    m = _FakeModel()
    ddp = GradientAllReduce(m)
    assert ddp.world == 1 and m.grad_ready_hook is None
    ddp.finish()
