"""Oracle restatement of the reference's losses and hard-label Dice metric (CPU, fp32/fp64).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def dice_terms(pred: Tensor, target: Tensor, dtype=torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """Per-(b,c) sums used by every overlap loss: I=sum p*t, P=sum p, T=sum t.

    reference: src/trainer/losses.py:49-69 (softmax over dim 1, one-hot target, flatten(2), sum(-1)).
    """
    p = torch.softmax(pred.detach().to("cpu", dtype), dim=1)
    C = p.shape[1]
    t = F.one_hot(target.cpu().long(), C).movedim(-1, 1).to(dtype)
    pf, tf = p.flatten(2), t.flatten(2)
    return (pf * tf).sum(-1), pf.sum(-1), tf.sum(-1)


def dice_loss(pred: Tensor, target: Tensor, smooth: float = 1.0, reduction: str = "mean",
              include_background: bool = True, dtype=torch.float32) -> Tensor:
    """DiceLoss.forward — src/trainer/losses.py:39-80."""
    I, P, T = dice_terms(pred, target, dtype)
    if not include_background:
        I, P, T = I[:, 1:], P[:, 1:], T[:, 1:]
    dl = 1.0 - (2.0 * I + smooth) / (P + T + smooth)
    if reduction == "mean":
        return dl.mean()
    if reduction == "sum":
        return dl.sum()
    return dl


def ce_loss(pred: Tensor, target: Tensor, class_weights: Optional[Tensor] = None, dtype=torch.float32) -> Tensor:
    """nn.CrossEntropyLoss(weight=class_weights) as used at losses.py:214,226 (mean reduction)."""
    w = None if class_weights is None else class_weights.to("cpu", dtype)
    return F.cross_entropy(pred.detach().to("cpu", dtype), target.cpu().long(), weight=w)


def dice_ce_loss(pred: Tensor, target: Tensor, dice_weight: float = 0.5, ce_weight: float = 0.5,
                 class_weights: Optional[Tensor] = None, include_background: bool = True,
                 dtype=torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """DiceCELoss.forward — src/trainer/losses.py:216-228.  Returns (total, dice_part, ce_part)."""
    d = dice_loss(pred, target, include_background=include_background, dtype=dtype)
    c = ce_loss(pred, target, class_weights, dtype)
    return dice_weight * d + ce_weight * c, d, c


def dice_ce_grad(pred: Tensor, target: Tensor, dice_weight: float = 0.5, ce_weight: float = 0.5,
                 dtype=torch.float64) -> Tensor:
    """d(DiceCE)/d(pred) by autograd on the restatement (used to check the fused backward kernel)."""
    z = pred.detach().to("cpu", dtype).requires_grad_(True)
    p = torch.softmax(z, dim=1)
    C = p.shape[1]
    t = F.one_hot(target.cpu().long(), C).movedim(-1, 1).to(dtype)
    pf, tf = p.flatten(2), t.flatten(2)
    I, U = (pf * tf).sum(-1), pf.sum(-1) + tf.sum(-1)
    d = (1.0 - (2.0 * I + 1.0) / (U + 1.0)).mean()
    c = F.cross_entropy(z, target.cpu().long())
    (dice_weight * d + ce_weight * c).backward()
    return z.grad


def focal_loss(pred: Tensor, target: Tensor, alpha: Optional[Tensor] = None, gamma: float = 2.0,
               dtype=torch.float32) -> Tensor:
    """FocalLoss.forward — src/trainer/losses.py:106-125 (mean reduction)."""
    w = None if alpha is None else alpha.to("cpu", dtype)
    ce = F.cross_entropy(pred.detach().to("cpu", dtype), target.cpu().long(), weight=w, reduction="none")
    pt = torch.exp(-ce)
    return ((1 - pt) ** gamma * ce).mean()


def tversky_loss(pred: Tensor, target: Tensor, alpha: float = 0.5, beta: float = 0.5, smooth: float = 1.0,
                 dtype=torch.float32) -> Tensor:
    """TverskyLoss.forward — src/trainer/losses.py:156-185 (TP=I, FP=P-I, FN=T-I)."""
    I, P, T = dice_terms(pred, target, dtype)
    tv = (I + smooth) / (I + alpha * (P - I) + beta * (T - I) + smooth)
    return (1.0 - tv).mean()


def dice_metric(pred_labels: Tensor, target: Tensor, num_classes: int,
                include_background: bool = False) -> Dict[str, object]:
    """DiceMetric.update + compute for one update — src/trainer/metrics.py:42-88 (hard labels, smooth 1e-5)."""
    pred_labels, target = pred_labels.cpu(), target.cpu()
    inter = torch.zeros(num_classes)
    union = torch.zeros(num_classes)
    for c in range(num_classes):
        pc, tc = (pred_labels == c), (target == c)
        inter[c] = (pc & tc).sum().float()
        union[c] = pc.sum().float() + tc.sum().float()
    dpc = (2.0 * inter + 1e-5) / (union + 1e-5)
    fg = dpc[0 if include_background else 1:]
    return {"dice": fg.mean().item(), "dice_per_class": dpc.tolist()}
