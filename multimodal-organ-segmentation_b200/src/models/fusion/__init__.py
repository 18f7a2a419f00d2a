"""Mirror of reference src/models/fusion/__init__.py."""
from .early_fusion import EarlyFusion
from .late_fusion import LateFusion, HierarchicalLateFusion
from .attention_fusion import AttentionFusion, CrossAttentionFusion, BidirectionalCrossAttention

__all__ = ["EarlyFusion", "LateFusion", "HierarchicalLateFusion", "AttentionFusion", "CrossAttentionFusion",
           "BidirectionalCrossAttention"]
