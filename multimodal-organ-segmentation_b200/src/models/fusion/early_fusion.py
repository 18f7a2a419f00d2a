"""Early fusion — drop-in mirror of the reference's src/models/fusion/early_fusion.py:13-64."""
from typing import Optional

import torch
import torch.nn as nn

from ..backbones.unet import _require_cuda, _no_autograd
from .... import kernels as K
from ....engine import ConvRunner
from ....kernels import Blocked


class EarlyFusion(nn.Module):
    def __init__(self, num_modalities: int = 2, in_channels_per_modality: int = 1, projection: bool = False,
                 out_channels: Optional[int] = None):
        super().__init__()
        self.num_modalities = num_modalities
        self.in_channels = num_modalities * in_channels_per_modality
        if projection:
            out_channels = out_channels or in_channels_per_modality
            self.proj = nn.Sequential(nn.Conv3d(self.in_channels, out_channels, kernel_size=1),
                                      nn.InstanceNorm3d(out_channels), nn.ReLU(inplace=True))
            self.out_channels = out_channels
        else:
            self.proj = nn.Identity()
            self.out_channels = self.in_channels
        self._runner = None

    def forward(self, x) -> torch.Tensor:
        if isinstance(x, (list, tuple)):
            x = torch.cat(x, dim=1)          # channel stacking of the raw inputs: a layout op, no arithmetic
        if isinstance(self.proj, nn.Identity):
            return x
        _require_cuda(x)
        _no_autograd(self, x)
        if self.out_channels % 16:
            raise NotImplementedError("EarlyFusion(projection=True) kernels need out_channels % 16 == 0")
        with torch.no_grad():
            B, C, Z, Y, X = x.shape
            src = Blocked(B, (C + 15) // 16 * 16, Z, Y, X, False, x.device)
            K.pack_ncdhw(x.contiguous().float(), src)
            if self._runner is None:
                self._runner = ConvRunner(False, x.device)
            pw = K.pack_conv_weight(self.proj[0].weight, None, False, [C], use_bias=False)
            out = Blocked(B, self.out_channels, Z, Y, X, False, x.device)
            self._runner.conv_norm_act(src, [(0, C)], pw, out)
            return out.to_ncdhw()
