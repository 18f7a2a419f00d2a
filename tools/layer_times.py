"""Per-launch times of one UNet3D forward (8 windows of 96^3, bf16) with every C-ABI call bracketed by CUDA events:
    python tools/layer_times.py [n_img] [mode]
Prints the conv layers sorted by time (the same table bench.py logs), for quick kernel experiments."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mmseg_b200  # noqa: F401
from mmseg_b200 import kernels as K
from mmseg_b200.src.models.backbones.unet import UNet3D

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(0)
m = UNet3D(in_channels=2, out_channels=8).eval().cuda().set_numeric_mode(mode)
x = torch.randn(n_img, 2, 96, 96, 96, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    K.PROFILE = []
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
prof, K.PROFILE = K.PROFILE, None
agg, layers = {}, {}
for name, info, a, b in prof:
    ms = a.elapsed_time(b)
    agg[name] = agg.get(name, 0.0) + ms / 3
    if info:
        L = layers.setdefault((name, info["layer"]), {"ms": 0.0, "n": 0, "info": info})
        L["ms"] += ms
        L["n"] += 1
print("totals:", {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}, "sum", round(sum(agg.values()), 3))
for (name, layer), L in sorted(layers.items(), key=lambda kv: -kv[1]["ms"]):
    i = L["info"]
    t = L["ms"] / L["n"]
    if name == "mmseg_conv3d_fwd":
        print(f"  {layer:36s} {t:7.3f} ms x{L['n'] // 3}  {i['flops'] / (t * 1e-3) / 1e12:7.1f} TF/s  tile {i['tile']} ctas {i['ctas']}")
    else:
        print(f"  norm {layer:31s} {t:7.3f} ms x{L['n'] // 3}  {i['bytes'] / (t * 1e-3) / 1e9:7.0f} GB/s")
