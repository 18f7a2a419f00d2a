"""Mirror of reference src/models/fusion/__init__.py (plus the module's other public classes)."""
from .early_fusion import EarlyFusion
from .late_fusion import LateFusion, HierarchicalLateFusion
from .attention_fusion import AttentionFusion, CrossAttentionFusion, BidirectionalCrossAttention, SUVGuidedAttention

__all__ = ["EarlyFusion", "LateFusion", "HierarchicalLateFusion", "AttentionFusion", "CrossAttentionFusion",
           "BidirectionalCrossAttention", "SUVGuidedAttention"]
