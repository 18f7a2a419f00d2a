"""Task heads with sm_100a kernels: SegmentationHead and DeepSupervisionHead.  (The reference's DetectionHead is never
instantiated by any of its builders and lies outside the hot path — SURVEY.md §2.1.)"""
from . import segmentation as _seg

SegmentationHead = _seg.SegmentationHead
DeepSupervisionHead = _seg.DeepSupervisionHead

__all__ = ("SegmentationHead", "DeepSupervisionHead")
