"""Late fusion — drop-in mirror of the reference's src/models/fusion/late_fusion.py (LateFusion :13-77,
HierarchicalLateFusion :80-135): concat (+ 1x1 conv, InstanceNorm, ReLU) / add / max / mean over modality features."""
from typing import List, Optional

import torch
import torch.nn as nn

from ..backbones.unet import _require_cuda, _no_autograd
from .... import kernels as K
from ....engine import ConvRunner
from ....kernels import Blocked


class LateFusion(nn.Module):
    def __init__(self, in_channels: int, num_modalities: int = 2, fusion_method: str = "concat",
                 out_channels: Optional[int] = None):
        super().__init__()
        self.in_channels = in_channels
        self.num_modalities = num_modalities
        self.fusion_method = fusion_method
        if fusion_method == "concat":
            self.out_channels = out_channels or in_channels
            self.proj = nn.Sequential(nn.Conv3d(in_channels * num_modalities, self.out_channels, kernel_size=1),
                                      nn.InstanceNorm3d(self.out_channels), nn.ReLU(inplace=True))
        else:
            self.out_channels = in_channels
            self.proj = nn.Identity()
        self._runner = None

    def forward(self, features: List[torch.Tensor]) -> torch.Tensor:
        _require_cuda(features[0])
        _no_autograd(self, features[0])
        B, C, Z, Y, X = features[0].shape
        M = len(features)
        if C % 16 or self.out_channels % 16:
            raise NotImplementedError("LateFusion kernels need channels % 16 == 0")
        with torch.no_grad():
            dev = features[0].device
            st = Blocked(B, M * C, Z, Y, X, False, dev)
            for m, f in enumerate(features):
                K.pack_ncdhw(f.contiguous().float(), st, c0=m * C)
            out = Blocked(B, self.out_channels, Z, Y, X, False, dev)
            method = self.fusion_method if self.fusion_method in ("add", "max", "mean") else "concat"
            if method == "concat":
                if not isinstance(self.proj, nn.Sequential):
                    return st.to_ncdhw()       # unknown method in the reference == plain concat without projection
                if self._runner is None:
                    self._runner = ConvRunner(False, dev)
                pw = K.pack_conv_weight(self.proj[0].weight, None, False, [C] * M, use_bias=False)
                self._runner.conv_norm_act(st, [(m * C, C) for m in range(M)], pw, out)
            elif method == "max":
                K.modality_max(st, M, C, out)
            else:
                K.modality_combine(st, M, C, out, 0, None, 1.0 if method == "add" else 1.0 / M)
            return out.to_ncdhw()


class HierarchicalLateFusion(nn.Module):
    def __init__(self, feature_channels: List[int], num_modalities: int = 2, fusion_method: str = "concat"):
        super().__init__()
        self.fusion_layers = nn.ModuleList(
            [LateFusion(in_channels=c, num_modalities=num_modalities, fusion_method=fusion_method) for c in feature_channels])

    def forward(self, multi_modal_features: List[List[torch.Tensor]]) -> List[torch.Tensor]:
        num_levels = len(multi_modal_features[0])
        return [self.fusion_layers[level]([modal[level] for modal in multi_modal_features]) for level in range(num_levels)]
