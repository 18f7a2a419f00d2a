"""Data-parallel gradient exchange on CPU: world_size 2 `gloo` groups run parallel.GradBucketReducer (both the
overlapped grad_ready/finish path and reduce_gradients) and must produce the mean of the per-rank gradients, including
with locally accumulated gradients from earlier micro-batches."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mmseg_b200  # noqa: F401
from mmseg_b200.parallel import GradBucketReducer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))


def _grads(rank, step):
    g = torch.Generator().manual_seed(100 * rank + step)
    return [torch.randn(p.shape, generator=g) for p in _model().parameters()]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _model()
        red = GradBucketReducer(m, bucket_bytes=64)   # tiny buckets -> several all-reduces
        assert len(red.buckets) > 1
        params = list(m.parameters())
        # micro-batch 0 accumulates locally; micro-batch 1 is armed and handed over in backward order
        for p, g in zip(params, _grads(rank, 0)):
            p.grad = g.clone()
        red.arm()
        for p, g in reversed(list(zip(params, _grads(rank, 1)))):
            red.grad_ready(p, g)
        red.finish()
        want = [torch.stack([_grads(r, 0)[i] + _grads(r, 1)[i] for r in range(world)]).mean(0) for i in range(len(params))]
        ok = all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(params, want))
        # plain path on populated .grad
        for p, g in zip(params, _grads(rank, 2)):
            p.grad = g.clone()
        red.reduce_gradients()
        want2 = [torch.stack([_grads(r, 2)[i] for r in range(world)]).mean(0) for i in range(len(params))]
        ok = ok and all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(params, want2))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
