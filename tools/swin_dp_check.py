"""Data-parallel SwinUNETR training check (2+ ranks, NCCL): the gradients the Trainer leaves after one data-parallel step
equal the mean of the per-rank gradients computed on one GPU.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/swin_dp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import mmseg_b200  # noqa: F401
from mmseg_b200.src.models.build import build_model
from mmseg_b200.src.trainer import Trainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = {"model": {"name": "swin_unetr", "in_channels": 2, "out_channels": 4, "backbone": {"feature_size": 48},
                 "fusion": {"type": "early"}, "head": {"dropout": 0.0}},
       "data": {"modalities": ["CT", "PET"]}, "hardware": {"device": "cuda", "mixed_precision": True},
       "training": {"epochs": 1, "accumulation_steps": 1, "optimizer": {"name": "adamw", "lr": 0.0, "weight_decay": 0.0},
                    "scheduler": {"name": "none"}, "loss": {"name": "dice_ce"}, "checkpoint": {}},
       "inference": {"batch_size": 1, "sliding_window": {"roi_size": [64, 64, 64], "overlap": 0.25}},
       "experiment": {"output_dir": "/tmp/swin_dp", "name": f"r{rank}"}}
torch.manual_seed(0)
model = build_model(cfg)


def batch(r):
    g = torch.Generator().manual_seed(100 + r)
    return torch.randn(1, 2, 64, 64, 64, generator=g).cuda(), torch.randint(0, 4, (1, 64, 64, 64), generator=g).cuda()


tr = Trainer(cfg, model, train_loader=[], val_loader=[])
assert tr.reducer is not None, "the Trainer did not set up the gradient reducer"
x, y = batch(rank)
outputs = tr.model(x)
loss = tr.criterion(outputs, y)
tr.reducer.arm()
loss.backward()
if any(tr.reducer._ready):
    tr.reducer.finish()
else:
    tr.reducer.reduce_gradients()
got = {n: p.grad.detach().clone() for n, p in tr.model.named_parameters()}
# single-GPU reference on every rank: mean over the ranks' batches
for p in tr.model.parameters():
    p.grad = None
ref = None
for r in range(world):
    xr, yr = batch(r)
    tr.criterion(tr.model(xr), yr).backward()
ref = {n: p.grad.detach() / world for n, p in tr.model.named_parameters()}
worst = max(((got[n] - ref[n]).norm() / ref[n].norm().clamp_min(1e-30)).item() for n in got)
t = torch.tensor([worst], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"[swin dp check] world {world}: worst relative difference of the all-reduced gradients vs the single-GPU mean: {t.item():.2e}")
    assert t.item() < 1e-5
dist.barrier()
dist.destroy_process_group()
