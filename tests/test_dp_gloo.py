"""N>1 host logic on CPU: world_size-2 gloo runs of the gradient bucketing / averaging layer (dl_vqa_b200/dp.py).

The stand-in models below mimic how VqaNet hands gradients over: ONE autograd node computes every parameter gradient,
fires `grad_ready_hook([(name, grad), ...])` stage by stage INSIDE backward and then returns the same tensors to
autograd -- so AccumulateGrad's steal / clone / accumulate behaviour is the real one."""
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


class _Node(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        ctx.model = model
        ctx.x = x
        return sum((p * x).sum() for p in params)

    @staticmethod
    def backward(ctx, g):
        m = ctx.model
        grads = m.make_grads(ctx.x, g)
        for stage in m.STAGES:
            if m.grad_ready_hook is not None:
                m.grad_ready_hook([(n, t) for n, t in grads.items() if n.startswith(stage + ".")])
        return (None, None) + tuple(grads[n] for n, _ in m.named_parameters())


class _Stage(torch.nn.Module):
    def __init__(self, shapes):
        super().__init__()
        for i, s in enumerate(shapes):
            self.register_parameter(f"w{i}", torch.nn.Parameter(torch.ones(*s)))


class _Model(torch.nn.Module):
    """Two stages; optional gradient arena with VqaNet's contract (views of a flat bucket per stage, fall back to fresh
    tensors while a live .grad still aliases the arena)."""
    STAGES = ("classifier", "text")

    def __init__(self):
        super().__init__()
        self.classifier = _Stage([(2, 3), (5,)])
        self.text = _Stage([(4,)])
        self.grad_ready_hook = None
        self._buckets = None

    def forward(self, x):
        return _Node.apply(self, x, *self.parameters())

    def make_grads(self, x, g):
        named = list(self.named_parameters())
        use = self._buckets is not None
        if use:
            owned = {b.untyped_storage().data_ptr() for b in self._buckets.values()}
            use = not any(p.grad is not None and p.grad.untyped_storage().data_ptr() in owned for _, p in named)
        out, off = {}, {s: 0 for s in self.STAGES}
        for n, p in named:
            val = torch.full_like(p, float(x) * float(g))
            if use:
                st = n.split(".")[0]
                view = self._buckets[st][off[st]:off[st] + p.numel()].view(p.shape)
                off[st] += p.numel()
                view.copy_(val)
                out[n] = view
            else:
                out[n] = val
        return out


def _use_gradient_arena(self, enable=True):
    self._buckets = ({s: torch.zeros(sum(p.numel() for p in getattr(self, s).parameters())) for s in self.STAGES}
                     if enable else None)
    return self


def _gradient_buckets(self):
    return self._buckets


class _ArenaModel(_Model):
    use_gradient_arena = _use_gradient_arena
    gradient_buckets = _gradient_buckets


class _PlainModel(_Model):
    """no arena methods on this class: dp.py sees none and uses the copying path"""


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _all_grads(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def _arena_worker(rank, world, port, out):
    _init(rank, world, port)
    from dl_vqa_b200.dp import GradientAllReduce
    m = _ArenaModel()
    ddp = GradientAllReduce(m)
    ok = ddp.sum_convention and ddp.grad_scale == 0.5 and m.gradient_buckets() is not None
    x = torch.tensor(float(rank + 1))                 # local gradient value: 1 on rank 0, 2 on rank 1
    # step 1: real backward through the hook, in place
    m(x).backward()
    ddp.finish()
    ok = ok and ddp.in_place and not ddp._pending
    ok = ok and torch.allclose(_all_grads(m), torch.full((15,), 3.0))            # SUM over ranks, in the arena
    flat = m.gradient_buckets()["classifier"]
    ok = ok and flat.data_ptr() <= m.classifier.w0.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4
    # step 2: a second backward WITHOUT zero_grad -> the model falls back to fresh tensors, autograd accumulates,
    # the wrapper must take the copying path and still leave (sum of step 1) + (sum of step 2) in p.grad
    m(x).backward()
    ddp.finish()
    ok = ok and (not ddp.in_place) and ddp.grad_scale == 0.5                      # scale convention never changes
    ok = ok and torch.allclose(_all_grads(m), torch.full((15,), 6.0))
    # step 3: after zero_grad(set_to_none=True) the in-place path is taken again (not latched off)
    for p in m.parameters():
        p.grad = None
    m(2 * x).backward()
    ddp.finish()
    ok = ok and ddp.in_place and torch.allclose(_all_grads(m), torch.full((15,), 6.0))
    # every rank holds identical gradients
    g = _all_grads(m).clone()
    others = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(others, g)
    ok = ok and all(torch.equal(o, g) for o in others)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_in_place_bucket_allreduce_through_real_backward_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_arena_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _plain_worker(rank, world, port, out):
    _init(rank, world, port)
    from dl_vqa_b200.dp import GradientAllReduce
    m = _PlainModel()
    with torch.no_grad():
        m.text.w0.fill_(float(rank + 7))
    ddp = GradientAllReduce(m)
    ok = m.grad_ready_hook is not None and not ddp.sum_convention and ddp.grad_scale == 1.0
    v0 = m.text.w0._version
    ddp.broadcast_parameters()
    ok = ok and float(m.text.w0[0]) == 7.0 and m.text.w0._version > v0           # rank 0's value, version bumped
    x = torch.tensor(float(rank + 1))
    m(x).backward()                                   # copying path: p.grad is NOT the tensor the hook saw being reduced
    ddp.finish()
    ok = ok and torch.allclose(_all_grads(m), torch.full((15,), 1.5)) and not ddp._pending     # averaged
    m(x).backward()                                   # accumulation without zero_grad
    ddp.finish()
    ok = ok and torch.allclose(_all_grads(m), torch.full((15,), 3.0))
    g = _all_grads(m).clone()
    others = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(others, g)
    ok = ok and all(torch.equal(o, g) for o in others)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_copying_allreduce_updates_p_grad_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_plain_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_single_process_is_a_noop():
    from dl_vqa_b200.dp import GradientAllReduce
    m = _PlainModel()
    ddp = GradientAllReduce(m)
    assert ddp.world == 1 and m.grad_ready_hook is None
    ddp.finish()
