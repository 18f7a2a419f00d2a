"""Per-launch CUDA-event times of one SwinUNETR-48 forward (eager, serialised): python tools/swin_layers.py [batch] [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmseg_b200  # noqa: F401
from mmseg_b200 import kernels as K
from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 96
torch.manual_seed(0)
m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).eval().cuda()
x = torch.randn(B, 2, S, S, S, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    K.PROFILE = []
    m(x)
    torch.cuda.synchronize()
    prof, K.PROFILE = K.PROFILE, None
tot = 0.0
for name, info, a, b in prof:
    ms = a.elapsed_time(b)
    tot += ms
    extra = ""
    if info:
        extra = info.get("layer", "")
        if "flops" in info:
            extra += f"  {info['flops'] / ms / 1e9:.0f} TF/s"
        if "bytes" in info:
            extra += f"  {info['bytes'] / ms / 1e6:.0f} GB/s"
        if "tile" in info:
            extra += f"  tile {info['tile']} ctas {info['ctas']}"
    print(f"{name[6:]:28s} {ms:7.3f} ms  {extra}")
print(f"total {tot:.3f} ms over {len(prof)} launches")
