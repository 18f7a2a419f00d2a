"""Public names of the training package of the sm_100a path — the three the reference's main.py imports
(`Trainer`, `get_loss`, `get_metrics`; reference src/trainer/__init__.py:5-7) plus the host-to-host inference helpers."""
from . import losses as _losses, metrics as _metrics, trainer as _trainer
from .inference import SlidingWindowInferer, predict_volume, sliding_window_inference

Trainer = _trainer.Trainer
get_loss = _losses.get_loss
get_metrics = _metrics.get_metrics

__all__ = ("Trainer", "get_loss", "get_metrics", "SlidingWindowInferer", "predict_volume", "sliding_window_inference")
