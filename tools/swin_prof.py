"""One SwinUNETR-48 training step at 2x96^3 for ncu: python tools/swin_prof.py [batch] [steps]
(ncu -k regex:swin_window_attention -c 16 ... captures the 8 forward + 8 backward window-attention launches of a step)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmseg_b200  # noqa: F401
from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
from mmseg_b200.src.trainer.losses import DiceCELoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).cuda().train()
crit = DiceCELoss()
x = torch.randn(B, 2, 96, 96, 96, device="cuda")
y = torch.randint(0, 8, (B, 96, 96, 96), device="cuda")
for _ in range(steps):
    loss = crit(m(x), y)
    loss.backward()
    m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print(f"SwinUNETR-48 B={B}: loss {loss.item():.5f}")
