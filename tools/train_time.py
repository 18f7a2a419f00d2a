"""Time the kernel training step: python tools/train_time.py [model unet|dual] [S] [B] [steps] [fusion] [n_modalities]
BASELINE.json configs[1]: dual 128 2 3 cross_attention 2;  configs[4]: dual 128 4 3 attention 4 (and unet 128 4 3 early 4)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmseg_b200
from mmseg_b200 import kernels as K
from mmseg_b200.src.models.build import build_model
from mmseg_b200.src.trainer.losses import DiceCELoss

kind = sys.argv[1] if len(sys.argv) > 1 else "dual"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
fusion = sys.argv[5] if len(sys.argv) > 5 else "cross_attention"
M = int(sys.argv[6]) if len(sys.argv) > 6 else 2
cfg = {"model": {"name": "dual_encoder" if kind == "dual" else "unet", "in_channels": M, "out_channels": 8,
                 "backbone": {"features": [32, 64, 128, 256, 512]}, "fusion": {"type": fusion},
                 "head": {"dropout": 0.1}},
       "data": {"modalities": ["CT", "PET", "MRI", "US"][:M]}, "hardware": {"device": "cuda"}}
torch.manual_seed(0)
m = build_model(cfg).train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5)
crit = DiceCELoss()
x = torch.randn(B, M, S, S, S, device="cuda")
y = torch.randint(0, 8, (B, S, S, S), device="cuda")
def step():
    loss = crit(m(x), y)
    loss.backward()
    opt.step(); opt.zero_grad()
    return loss
for _ in range(2):
    l = step()
torch.cuda.synchronize()
print("warm loss", l.item(), "mem GB", torch.cuda.max_memory_allocated() / 2**30)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for _ in range(steps):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{kind} S={S} B={B}: {ms:.2f} ms/step (host {1e3 * (time.time() - t0) / steps:.2f} ms) -> {B / ms * 1e3:.2f} samples/s; loss {l.item():.5f}")
os.environ["MMSEG_WGRAD_SIDE_STREAM"] = "0"      # serialised: per-kernel times without overlap
step()
torch.cuda.synchronize()
K.PROFILE = []
step()
torch.cuda.synchronize()
prof, K.PROFILE = K.PROFILE, None
agg = {}
for name, info, a, b in prof:
    key = name if not info or "layer" not in info else name + ":" + (info["layer"] if "wgrad" in info["layer"] or "conv3d_fwd" in name else info["layer"].split(" ")[0])
    d = agg.setdefault(key, [0.0, 0, 0.0, 0.0])
    d[0] += a.elapsed_time(b); d[1] += 1
    if info: d[2] += info.get("flops", 0.0); d[3] += info.get("bytes", 0.0) if "flops" not in info else 0.0
tot = sum(d[0] for d in agg.values())
print(f"kernel time inside C-ABI calls: {tot:.2f} ms")
for k_, d in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    tf = f"{d[2] / d[0] / 1e9:8.1f} TF/s" if d[2] else (f"{d[3] / d[0] / 1e6:8.0f} GB/s" if d[3] else "")
    print(f"  {k_:44s} {d[0]:8.3f} ms {100 * d[0] / tot:5.1f}%  n={d[1]:3d} {tf}")
