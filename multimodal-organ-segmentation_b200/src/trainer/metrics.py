"""Evaluation metrics — drop-in mirror of the reference's src/trainer/metrics.py for the hot path (DiceMetric,
ConfusionMatrix, get_metrics).  Predictions and labels stay on the GPU: one pass builds the K x K confusion counts
(csrc/dicece.cu confusion_kernel) from which both metrics are read; the reference copies every batch to the CPU and,
for ConfusionMatrix, loops over voxels in Python (metrics.py:195-196).  HausdorffDistance (scipy EDT) is not on the path.
"""
from typing import Any, Dict, List

import torch

from ... import kernels as K


class _ConfusionState:
    def __init__(self, num_classes: int):
        self.num_classes = num_classes
        self.counts = None

    def reset(self):
        self.counts = None

    def add(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        if not pred.is_cuda:
            raise RuntimeError("mmseg_b200 metrics run on CUDA tensors only (no CPU fallback)")
        if self.counts is None:
            self.counts = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=pred.device)
        K.confusion_hist(pred, target.to(pred.device), self.num_classes, self.counts)


class DiceMetric:
    """reference metrics.py:11-88: hard-label Dice accumulated over all update() calls, smooth 1e-5."""

    def __init__(self, num_classes: int, include_background: bool = False, reduction: str = "mean"):
        self.num_classes = num_classes
        self.include_background = include_background
        self.reduction = reduction
        self._state = _ConfusionState(num_classes)
        self.reset()

    def reset(self) -> None:
        self._state.reset()
        self.count = 0

    def update(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        self._state.add(pred, target)
        self.count += 1

    @property
    def intersection(self) -> torch.Tensor:
        return torch.diagonal(self._state.counts).float().cpu()

    @property
    def union(self) -> torch.Tensor:
        c = self._state.counts
        return (c.sum(0) + c.sum(1)).float().cpu()

    def compute(self) -> Dict[str, Any]:
        smooth = 1e-5
        dice_per_class = (2.0 * self.intersection + smooth) / (self.union + smooth)
        start_idx = 0 if self.include_background else 1
        return {"dice": dice_per_class[start_idx:].mean().item(), "dice_per_class": dice_per_class.tolist()}


class ConfusionMatrix:
    """reference metrics.py:165-226 (rows = target, columns = prediction)."""

    def __init__(self, num_classes: int):
        self.num_classes = num_classes
        self._state = _ConfusionState(num_classes)

    def reset(self) -> None:
        self._state.reset()

    def update(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        self._state.add(pred, target)

    @property
    def matrix(self) -> torch.Tensor:
        return self._state.counts.cpu()

    def compute(self) -> Dict[str, Any]:
        m = self.matrix.double()
        tp = torch.diagonal(m)
        fp = m.sum(0) - tp
        fn = m.sum(1) - tp
        precision = tp / (tp + fp + 1e-8)
        recall = tp / (tp + fn + 1e-8)
        f1 = 2 * precision * recall / (precision + recall + 1e-8)
        accuracy = tp.sum() / (m.sum() + 1e-8)
        return {"accuracy": accuracy.item(), "precision": precision.mean().item(), "recall": recall.mean().item(),
                "f1": f1.mean().item(), "precision_per_class": precision.tolist(), "recall_per_class": recall.tolist(),
                "f1_per_class": f1.tolist(), "confusion_matrix": self.matrix.tolist()}


def get_metrics(config: Dict[str, Any]) -> Dict[str, Any]:
    """reference metrics.py:229-244."""
    num_classes = config["model"]["out_channels"]
    return {"dice": DiceMetric(num_classes=num_classes), "confusion": ConfusionMatrix(num_classes=num_classes)}
