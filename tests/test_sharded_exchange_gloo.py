"""N>1 path of sliding-window inference on CPU: world_size 2 and 3 `gloo` process groups run the product's exchange
plan / partial exchange / label gather on synthetic partial accumulators and must reproduce the single-rank sums."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mmseg_b200  # noqa: F401
from mmseg_b200.src.trainer import inference as INF


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial(starts, lo, hi, roi, vol, K):
    """What a rank's accumulate() produces for windows [lo, hi): a deterministic integer-valued 'logit' per window
    (exact in fp32, so sums are order independent and equality can be bit-exact)."""
    acc = torch.zeros((K + 1, *vol))
    for i in range(lo, hi):
        z, y, x = starts[i]
        sl = (slice(z, z + roi[0]), slice(y, y + roi[1]), slice(x, x + roi[2]))
        for c in range(K):
            acc[(c,) + sl] += float((i * 7 + c * 3) % 11)
        acc[(K,) + sl] += 1.0
    return acc


def _worker(rank, world, port, vol, roi, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        starts = INF.window_starts(vol, roi, 0.5)
        lo, hi = INF.shard_windows(len(starts), world, rank)
        acc = _partial(starts, lo, hi, roi, vol, K)
        plan = INF.exchange_plan(starts, roi[0], vol[0], world)
        INF.exchange_partials(acc, plan, rank, world)
        z0, z1 = plan["slabs"][rank]
        full = _partial(starts, 0, len(starts), roi, vol, K)
        ok = torch.equal(acc[:, z0:z1], full[:, z0:z1])
        lab = (acc[:K, z0:z1] / acc[K, z0:z1]).argmax(0).to(torch.uint8) if z1 > z0 else None
        labels = INF.gather_label_slabs(lab, plan["slabs"], vol, rank, world, None, torch.device("cpu"))
        want = (full[:K] / full[K]).argmax(0).to(torch.uint8)
        ok = ok and torch.equal(labels, want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,vol", [(2, (40, 20, 24)), (3, (56, 16, 20)), (2, (16, 16, 40))])
def test_exchange_and_gather(world, vol):
    roi, K = (16, 16, 16), 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, vol, roi, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)], res


def test_exchange_plan_covers_every_foreign_contribution():
    starts = INF.window_starts((512, 512, 300), (96, 96, 96), 0.5)
    for world in (2, 4, 8):
        plan = INF.exchange_plan(starts, 96, 512, world)
        for src in range(world):
            t0, t1 = plan["touched"][src]
            covered = []
            for dst in range(world):
                rng = plan["slabs"][dst] if dst == src else plan["sends"][src][dst]
                if rng is None:
                    continue
                a, b = max(rng[0], t0), min(rng[1], t1)
                if b > a:
                    covered.append((a, b))
            covered.sort()
            assert covered[0][0] == t0 and covered[-1][1] == t1
            assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
