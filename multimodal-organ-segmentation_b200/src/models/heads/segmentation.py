"""Segmentation head — drop-in mirror of the reference's src/models/heads/segmentation.py:11-59 (SegmentationHead):
Dropout3d -> Conv3d(k, padding k//2) -> {softmax, sigmoid, identity}.  The conv runs in the tcgen05 kernel and writes
NCDHW fp32 logits (k = 1 heads with <= 16 classes on the CUDA-core 1x1 kernel); the optional channel softmax / sigmoid
is applied on those logits.  DeepSupervisionHead (:62-115): one head per scale + trilinear (align_corners=True) resize of
the coarse logits to target_size in a native kernel."""
from typing import Optional

import torch
import torch.nn as nn

from ..backbones.unet import _require_cuda, _no_autograd
from .... import kernels as K
from .... import _lib
from ....kernels import Blocked


class SegmentationHead(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 1, dropout: float = 0.0,
                 activation: Optional[str] = None):
        super().__init__()
        self.dropout = nn.Dropout3d(dropout) if dropout > 0 else nn.Identity()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, padding=kernel_size // 2)
        if activation == "softmax":
            self.activation = nn.Softmax(dim=1)
        elif activation == "sigmoid":
            self.activation = nn.Sigmoid()
        else:
            self.activation = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x)
        _no_autograd(self, x)
        if self.conv.kernel_size[0] not in (1, 3):
            raise NotImplementedError("SegmentationHead kernels cover kernel_size 1 and 3")
        if self.training and isinstance(self.dropout, nn.Dropout3d):
            raise NotImplementedError("train-mode Dropout3d in a stand-alone head is not built; call .eval()")
        with torch.no_grad():
            B, C, Z, Y, X = x.shape
            src = Blocked(B, (C + 15) // 16 * 16, Z, Y, X, False, x.device)
            K.pack_ncdhw(x.contiguous().float(), src)
            out = torch.empty((B, self.conv.out_channels, Z, Y, X), dtype=torch.float32, device=x.device)
            if self.conv.kernel_size[0] == 1 and C % 8 == 0 and C <= 256 and self.conv.out_channels <= 16:
                K.conv1x1_logits(src, 0, C, self.conv.weight, self.conv.bias, out)
            else:
                pw = K.pack_conv_weight(self.conv.weight, self.conv.bias, False, [C])
                K.conv3d(src, pw, K.a_chunk_table(src, [0], [C], False), out, _lib.OUT_NCDHW_F32)
            return self.activation(out)   # softmax / sigmoid over 8 channels: epilogue-sized, left to torch


class DeepSupervisionHead(nn.Module):
    """Mirror of segmentation.py:62-115: a SegmentationHead per scale; predictions whose spatial size differs from
    target_size are resized with trilinear interpolation, align_corners=True (kernels.trilinear_resize)."""

    def __init__(self, in_channels_list: list, out_channels: int, dropout: float = 0.0):
        super().__init__()
        self.heads = nn.ModuleList(SegmentationHead(c, out_channels, dropout=dropout) for c in in_channels_list)

    def forward(self, features: list, target_size: Optional[tuple] = None) -> list:
        outputs = []
        for feat, head in zip(features, self.heads):
            out = head(feat)
            if target_size is not None and tuple(out.shape[2:]) != tuple(target_size):
                out = K.trilinear_resize(out, target_size)
            outputs.append(out)
        return outputs
