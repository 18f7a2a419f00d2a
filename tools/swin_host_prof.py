"""Host-side (Python) profile of the SwinUNETR training step — the step is launch-bound at B <= 2:
python tools/swin_host_prof.py [batch]   -> cProfile of 3 steps, top functions by own time and by cumulative time."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmseg_b200  # noqa: F401
from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
from mmseg_b200.src.trainer.losses import DiceCELoss
from mmseg_b200.optim import FusedAdamW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).cuda().train()
opt = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=1e-5)
crit = DiceCELoss()
x = torch.randn(B, 2, 96, 96, 96, device="cuda")
y = torch.randint(0, 8, (B, 96, 96, 96), device="cuda")


def step():
    loss = crit(m(x), y)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t_issue = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5
print(f"B={B}: host issue time {1e3 * t_issue:.2f} ms/step, with the final sync {1e3 * t_all:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
