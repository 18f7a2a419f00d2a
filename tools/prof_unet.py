"""Profiling target: N forwards of the full UNet3D (96^3 windows, bf16 mode) through the engine.
    python tools/prof_unet.py [n_img] [iters] [mode]
One forward = 1 pack + 18 x (conv, finalize, apply) + 4 convT + 1 logits conv = 60 launches, 23 of them conv3d_tc_kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mmseg_b200  # noqa: F401
from mmseg_b200.src.models.backbones.unet import UNet3D

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
m = UNet3D(in_channels=2, out_channels=8).eval().cuda().set_numeric_mode(mode)
x = torch.randn(n_img, 2, 96, 96, 96, device="cuda")
with torch.no_grad():
    for _ in range(iters):
        y = m(x)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
