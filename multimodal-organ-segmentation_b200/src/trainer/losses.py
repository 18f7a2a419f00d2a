"""Loss functions — drop-in mirror of the reference's src/trainer/losses.py (same class names, constructor arguments,
`forward(pred[B,C,H,W,D] float, target[B,H,W,D] int64) -> scalar`, and `get_loss(config)` factory).

DiceLoss / CrossEntropy / DiceCELoss run in the one-pass sm_100a kernel (csrc/dicece.cu): logits and labels are read
once, softmax in registers, per-(batch, class) reductions, deterministic finalize; the backward kernel recomputes the
softmax and writes d(logits).  No PyTorch fallback: CPU tensors raise.
"""
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from ... import kernels as K


class _DiceCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, dice_weight, ce_weight, smooth, include_background, class_weights):
        if not pred.is_cuda:
            raise RuntimeError("mmseg_b200 losses run on CUDA tensors only (no CPU fallback)")
        logits = pred.detach().contiguous().float()
        tgt = target.detach().contiguous().long()
        cw = None if class_weights is None else class_weights.detach().to(pred.device, torch.float32).contiguous()
        result, sums = K.dicece_fwd(logits, tgt, dice_weight, ce_weight, smooth, include_background, cw)
        ctx.save_for_backward(logits, tgt, sums)
        ctx.cfg = (dice_weight, ce_weight, smooth, include_background, cw)
        ctx.in_dtype = pred.dtype
        return result[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        logits, tgt, sums = ctx.saved_tensors
        dw, cwt, smooth, bg, cw = ctx.cfg
        dl = K.dicece_bwd(logits, tgt, sums, grad_out, dw, cwt, smooth, bg, cw)
        return dl.to(ctx.in_dtype), None, None, None, None, None, None


class DiceLoss(nn.Module):
    """reference losses.py:12-80."""

    def __init__(self, smooth: float = 1.0, reduction: str = "mean", softmax: bool = True, include_background: bool = True):
        super().__init__()
        if not softmax:
            raise NotImplementedError("DiceLoss(softmax=False) (predictions that are already probabilities) has no "
                                      "sm_100a kernel: the one-pass kernel applies the softmax itself")
        self.smooth = smooth
        self.reduction = reduction
        self.softmax = softmax
        self.include_background = include_background

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("DiceLoss(reduction='none') is not built in the sm_100a path (mean | sum)")
        loss = _DiceCEFunction.apply(pred, target, 1.0, 0.0, self.smooth, self.include_background, None)
        if self.reduction == "sum":
            n_cls = pred.shape[1] - (0 if self.include_background else 1)
            loss = loss * (pred.shape[0] * n_cls)
        return loss


class _FocalFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, alpha, gamma):
        if not pred.is_cuda:
            raise RuntimeError("mmseg_b200 losses run on CUDA tensors only (no CPU fallback)")
        logits = pred.detach().contiguous().float()
        tgt = target.detach().contiguous().long()
        cw = None if alpha is None else alpha.detach().to(pred.device, torch.float32).contiguous()
        ctx.save_for_backward(logits, tgt)
        ctx.cfg, ctx.in_dtype = (cw, gamma), pred.dtype
        return K.focal(logits, tgt, cw, gamma)[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        logits, tgt = ctx.saved_tensors
        cw, gamma = ctx.cfg
        return K.focal(logits, tgt, cw, gamma, grad_out, backward=True).to(ctx.in_dtype), None, None, None


class FocalLoss(nn.Module):
    """reference losses.py:83-125 (one-pass sm_100a kernel; reduction mean | sum)."""

    def __init__(self, alpha: Optional[torch.Tensor] = None, gamma: float = 2.0, reduction: str = "mean"):
        super().__init__()
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("FocalLoss(reduction='none') is not built in the sm_100a path (mean | sum)")
        loss = _FocalFunction.apply(pred, target, self.alpha, self.gamma)
        return loss * target.numel() if self.reduction == "sum" else loss


class _TverskyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, alpha, beta, smooth):
        if not pred.is_cuda:
            raise RuntimeError("mmseg_b200 losses run on CUDA tensors only (no CPU fallback)")
        logits = pred.detach().contiguous().float()
        tgt = target.detach().contiguous().long()
        result, sums = K.tversky_fwd(logits, tgt, alpha, beta, smooth)
        ctx.save_for_backward(logits, tgt, sums)
        ctx.cfg, ctx.in_dtype = (alpha, beta, smooth), pred.dtype
        return result[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        logits, tgt, sums = ctx.saved_tensors
        a, b, s = ctx.cfg
        return K.tversky_bwd(logits, tgt, sums, grad_out, a, b, s).to(ctx.in_dtype), None, None, None, None


class TverskyLoss(nn.Module):
    """reference losses.py:128-185 (same one-pass sums as Dice: TP = I, FP = P - I, FN = T - I; reduction mean | sum)."""

    def __init__(self, alpha: float = 0.5, beta: float = 0.5, smooth: float = 1.0, reduction: str = "mean"):
        super().__init__()
        self.alpha, self.beta, self.smooth, self.reduction = alpha, beta, smooth, reduction

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("TverskyLoss(reduction='none') is not built in the sm_100a path (mean | sum)")
        loss = _TverskyFunction.apply(pred, target, self.alpha, self.beta, self.smooth)
        return loss * (pred.shape[0] * pred.shape[1]) if self.reduction == "sum" else loss


class _CrossEntropy(nn.Module):
    """nn.CrossEntropyLoss(weight=class_weights) (reference losses.py:248-249) through the same one-pass kernel."""

    def __init__(self, weight: Optional[torch.Tensor] = None):
        super().__init__()
        self.register_buffer("weight", weight)

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _DiceCEFunction.apply(pred, target, 0.0, 1.0, 1.0, True, self.weight)


class DiceCELoss(nn.Module):
    """reference losses.py:188-228: dice_weight * DiceLoss + ce_weight * CrossEntropyLoss(weight=class_weights)."""

    def __init__(self, dice_weight: float = 0.5, ce_weight: float = 0.5, class_weights: Optional[torch.Tensor] = None,
                 include_background: bool = True):
        super().__init__()
        self.dice_weight = dice_weight
        self.ce_weight = ce_weight
        self.dice_loss = DiceLoss(include_background=include_background)
        self.ce_loss = _CrossEntropy(class_weights)

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _DiceCEFunction.apply(pred, target, self.dice_weight, self.ce_weight, self.dice_loss.smooth,
                                     self.dice_loss.include_background, self.ce_loss.weight)


def get_loss(config: Dict[str, Any]) -> nn.Module:
    """reference losses.py:231-267."""
    loss_config = config["training"]["loss"]
    loss_name = loss_config["name"].lower()
    class_weights = loss_config.get("class_weights")
    if class_weights is not None:
        class_weights = torch.tensor(class_weights, dtype=torch.float32)
    if loss_name == "dice":
        return DiceLoss()
    elif loss_name == "ce" or loss_name == "cross_entropy":
        return _CrossEntropy(class_weights)
    elif loss_name == "dice_ce":
        return DiceCELoss(dice_weight=loss_config.get("dice_weight", 0.5), ce_weight=loss_config.get("ce_weight", 0.5),
                          class_weights=class_weights)
    elif loss_name == "focal":
        return FocalLoss(alpha=class_weights)
    elif loss_name == "tversky":
        return TverskyLoss(alpha=loss_config.get("tversky_alpha", 0.5), beta=loss_config.get("tversky_beta", 0.5))
    else:
        return DiceCELoss()
