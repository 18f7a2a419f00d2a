"""Mirror of reference src/trainer/__init__.py:5-7."""
from .trainer import Trainer
from .losses import get_loss
from .metrics import get_metrics

__all__ = ["Trainer", "get_loss", "get_metrics"]
