"""Dual encoder — placeholder until the fused engine lands (filled in below in the same round)."""
from typing import Any, Dict

import torch.nn as nn


class DualEncoder(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("DualEncoder engine not built yet")


def build_dual_encoder(config: Dict[str, Any]) -> DualEncoder:
    return DualEncoder()
