"""Whole-model gradient check of SwinUNETR with every parameter's relative L2 error: python tools/swin_train_check.py [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mmseg_b200  # noqa: F401
from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
from mmseg_b200.src.trainer.losses import DiceCELoss
from oracle import swin_unetr as O

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).train()
with torch.no_grad():
    for name, p in m.named_parameters():
        if "relative_position_bias_table" in name:
            p.normal_(0, 0.3)
sd64 = {k: v.detach().double().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
x = torch.randn(1, 2, size, size, size)
y = torch.randint(0, 8, (1, size, size, size))
O._p = lambda sd, key, dtype: sd[key]
ref_logits = O.swin_unetr_forward(sd64, x, dtype=torch.float64)
p_ = torch.softmax(ref_logits, dim=1).flatten(2)
t_ = F.one_hot(y, 8).movedim(-1, 1).double().flatten(2)
dice = (1.0 - (2.0 * (p_ * t_).sum(-1) + 1.0) / (p_.sum(-1) + t_.sum(-1) + 1.0)).mean()
ref_loss = 0.5 * dice + 0.5 * F.cross_entropy(ref_logits, y)
ref_loss.backward()
m = m.cuda()
logits = m(x.cuda())
loss = DiceCELoss()(logits, y.cuda())
loss.backward()
print(f"loss {loss.item():.6f} vs {ref_loss.item():.6f}; logits rel {((logits.detach().cpu().double() - ref_logits.detach()).norm() / ref_logits.detach().norm()).item():.2e}")
for name, p in m.named_parameters():
    g, r = p.grad.cpu().double(), sd64[name].grad
    e = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
    print(f"{e:9.2e}  |g| {g.norm().item():9.2e} |ref| {r.norm().item():9.2e}  {name}")

# the reference arithmetic itself under bf16 autocast (what `use_amp` would do to the reference): the floor of this comparison
sd32 = {k: v.detach().float().clone().requires_grad_(v.is_floating_point()) for k, v in sd64.items()}
with torch.autocast("cpu", dtype=torch.bfloat16):
    lg = O.swin_unetr_forward(sd32, x, dtype=torch.float32)
lg = lg.float()
p_ = torch.softmax(lg, dim=1).flatten(2)
t32 = t_.float()
dice = (1.0 - (2.0 * (p_ * t32).sum(-1) + 1.0) / (p_.sum(-1) + t32.sum(-1) + 1.0)).mean()
(0.5 * dice + 0.5 * F.cross_entropy(lg, y)).backward()
ours, auto = [], []
for name, p in m.named_parameters():
    r = sd64[name].grad
    ours.append(((p.grad.cpu().double() - r).norm() / r.norm().clamp_min(1e-30)).item())
    auto.append(((sd32[name].grad.double() - r).norm() / r.norm().clamp_min(1e-30)).item())
ours_s, auto_s = sorted(ours), sorted(auto)
print(f"SUMMARY kernels vs fp64: median {ours_s[len(ours_s) // 2]:.3f} worst {ours_s[-1]:.3f} | reference under bf16 autocast vs fp64: "
      f"median {auto_s[len(auto_s) // 2]:.3f} worst {auto_s[-1]:.3f}; logits rel (autocast) "
      f"{((lg.detach().double() - ref_logits.detach()).norm() / ref_logits.detach().norm()).item():.2e}")
