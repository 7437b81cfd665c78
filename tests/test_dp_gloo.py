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


class _ArenaModel(torch.nn.Module):
    """Mimics VqaNet.use_gradient_arena(): the gradients of a stage are views of one flat bucket."""
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(2, 3))
        self.grad_ready_hook = None
        self._buckets = None

    def use_gradient_arena(self, enable=True):
        self._buckets = {"classifier": torch.zeros(12)} if enable else None
        return self

    def gradient_buckets(self):
        return self._buckets


def _arena_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dl_vqa_b200.dp import GradientAllReduce
    m = _ArenaModel()
    ddp = GradientAllReduce(m)
    flat = m.gradient_buckets()["classifier"]
    assert ddp.in_place and flat is not None
    a, b = flat[0:6].view(2, 3), flat[8:12]
    a.fill_(float(rank + 1)); b.fill_(10.0 * (rank + 1))
    ddp._on_group_ready([("classifier.a", a), ("classifier.b", b)])
    ddp.finish()
    # summed in place, averaging deferred to the optimizer through grad_scale
    ok = (torch.allclose(a, torch.full((2, 3), 3.0)) and torch.allclose(b, torch.full((4,), 30.0))
          and ddp.grad_scale == 0.5 and not ddp._pending)
    # gradients that are NOT arena views fall back to the copying path (and finish() averages them)
    g = [("classifier.x", torch.full((3,), float(rank)))]
    ddp._on_group_ready(g)
    ddp.finish()
    ok = ok and torch.allclose(g[0][1], torch.full((3,), 0.5)) and ddp.grad_scale == 1.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_in_place_bucket_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_arena_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


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
