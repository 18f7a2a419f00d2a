"""Fusion modules of the sm_100a path under the reference's class names (reference src/models/fusion/__init__.py exports
the first four; the remaining classes of its modules are exported here as well)."""
from . import attention_fusion as _af, early_fusion as _ef, late_fusion as _lf

EarlyFusion = _ef.EarlyFusion
LateFusion, HierarchicalLateFusion = _lf.LateFusion, _lf.HierarchicalLateFusion
AttentionFusion, CrossAttentionFusion = _af.AttentionFusion, _af.CrossAttentionFusion
BidirectionalCrossAttention, SUVGuidedAttention = _af.BidirectionalCrossAttention, _af.SUVGuidedAttention

__all__ = ("EarlyFusion", "LateFusion", "HierarchicalLateFusion", "AttentionFusion", "CrossAttentionFusion",
           "BidirectionalCrossAttention", "SUVGuidedAttention")
