"""Mirror of reference src/models/heads/__init__.py (DetectionHead is out of scope: never instantiated)."""
from .segmentation import SegmentationHead, DeepSupervisionHead

__all__ = ["SegmentationHead", "DeepSupervisionHead"]
