"""Dual encoder — drop-in mirror of the reference's src/models/backbones/dual_encoder.py.

Same constructor arguments, attribute tree (`encoders[i]["init_conv"|"blocks"]`, `fusion_layers` | `fusion_proj`,
`decoder`, `dropout`, `out_conv`) and state_dict keys/shapes; the arithmetic runs in the sm_100a
kernels (engine.DualEncoderEngine).  No PyTorch/CPU route.
"""
from typing import Any, Dict, List, Tuple, Union

import torch
import torch.nn as nn

from .unet import (ConvBlock3D, DownBlock3D, UpBlock3D, _require_cuda, _no_autograd, _wants_grad, _train_step_forward,
                   _DEFAULT_MODE)
from ....engine import DualEncoderEngine
from .... import kernels as K
from ....kernels import Blocked
from ....numerics import mode as numeric_mode


class CrossModalAttention(nn.Module):
    """SE-style modality gate — reference dual_encoder.py:207-254 (same Sequential indices: Linear at 2 and 4)."""

    def __init__(self, channels: int, num_modalities: int, reduction: int = 4):
        super().__init__()
        self.channels = channels
        self.num_modalities = num_modalities
        self.attention = nn.Sequential(
            nn.AdaptiveAvgPool3d(1),
            nn.Flatten(),
            nn.Linear(channels * num_modalities, channels * num_modalities // reduction),
            nn.ReLU(inplace=True),
            nn.Linear(channels * num_modalities // reduction, num_modalities),
            nn.Softmax(dim=1),
        )
        self.numeric_mode = _DEFAULT_MODE

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, M, C, H, W, D] -> [B, C, H, W, D] (stand-alone use; DualEncoder runs the same kernels in its engine)."""
        _require_cuda(x)
        _no_autograd(self, x)
        B, M, C, H, W, D = x.shape
        if C % 16:
            raise NotImplementedError("CrossModalAttention kernels need channels % 16 == 0")
        split = numeric_mode(self.numeric_mode)
        with torch.no_grad():
            st = Blocked(B, M * C, H, W, D, split, x.device)
            K.pack_ncdhw(x.reshape(B, M * C, H, W, D).contiguous().float(), st)
            pooled = K.channel_mean(st, 0, M * C)
            att = self.attention
            w = K.gate_mlp(pooled, att[2].weight, att[2].bias, att[4].weight, att[4].bias)
            out = Blocked(B, C, H, W, D, split, x.device)
            K.modality_combine(st, M, C, out, 0, w)
            return out.to_ncdhw()


class DualEncoder(nn.Module):
    """reference dual_encoder.py:15-204."""

    def __init__(self, in_channels_per_modality: int = 1, num_modalities: int = 2, out_channels: int = 8,
                 features: List[int] = [32, 64, 128, 256, 512], norm: str = "instance", fusion_type: str = "concat",
                 dropout: float = 0.0, shared_decoder: bool = True, **kwargs):
        super().__init__()
        self.in_channels_per_modality = in_channels_per_modality
        self.num_modalities = num_modalities
        self.out_channels = out_channels
        self.features = list(features)
        self.fusion_type = fusion_type
        self.shared_decoder = shared_decoder
        self.encoders = nn.ModuleList()
        for _ in range(num_modalities):
            self.encoders.append(self._build_encoder(in_channels_per_modality, self.features, norm))
        if fusion_type == "attention":
            self.fusion_layers = nn.ModuleList([CrossModalAttention(feat, num_modalities) for feat in self.features])
        elif fusion_type == "concat":
            self.fusion_proj = nn.ModuleList(
                [nn.Conv3d(feat * num_modalities, feat, kernel_size=1) for feat in self.features])
        self.decoder = self._build_decoder(self.features, norm)
        self.dropout = nn.Dropout3d(dropout) if dropout > 0 else nn.Identity()
        self.out_conv = nn.Conv3d(self.features[0], out_channels, kernel_size=1)
        self.numeric_mode = _DEFAULT_MODE
        self._engines: Dict[str, DualEncoderEngine] = {}

    def _build_encoder(self, in_channels: int, features: List[int], norm: str) -> nn.ModuleDict:
        encoder = nn.ModuleDict()
        encoder["init_conv"] = ConvBlock3D(in_channels, features[0], norm=norm)
        encoder["blocks"] = nn.ModuleList()
        for i in range(len(features) - 1):
            encoder["blocks"].append(DownBlock3D(features[i], features[i + 1], norm=norm))
        return encoder

    def _build_decoder(self, features: List[int], norm: str) -> nn.ModuleList:
        decoder = nn.ModuleList()
        for i in range(len(features) - 1, 0, -1):
            decoder.append(UpBlock3D(features[i], features[i - 1], norm=norm))
        return decoder

    def set_numeric_mode(self, mode: str) -> "DualEncoder":
        self.numeric_mode = numeric_mode(mode).name
        return self

    def engine(self) -> DualEncoderEngine:
        e = self._engines.get(self.numeric_mode)
        if e is None:
            e = self._engines[self.numeric_mode] = DualEncoderEngine(self, self.numeric_mode)
        return e

    def forward(self, x: torch.Tensor, return_features: bool = False
                ) -> Union[torch.Tensor, Tuple[torch.Tensor, Dict[str, List[torch.Tensor]]]]:
        _require_cuda(x)
        self.encoders[0]["init_conv"].kernel_supported()
        if _wants_grad(self, x):
            if return_features:
                raise NotImplementedError("return_features is an inference-path option")
            return _train_step_forward(self, "dual", x)
        if self.training and isinstance(self.dropout, nn.Dropout3d) and self.dropout.p > 0:
            raise NotImplementedError("train-mode Dropout3d under no_grad: call model.eval() for inference")
        eng = self.engine()
        logits = eng.forward(x)
        if return_features:
            n, _, Z, Y, X = x.shape
            enc, fused = eng.features_ncdhw(n, Z, Y, X)
            return logits, {"encoder_features": enc, "fused_features": fused}
        return logits

    @property
    def encoder_channels(self) -> List[int]:
        return self.features


def build_dual_encoder(config: Dict[str, Any]) -> DualEncoder:
    """reference dual_encoder.py:257-280."""
    backbone_config = config.get("model", {}).get("backbone", {})
    fusion_config = config.get("model", {}).get("fusion", {})
    num_modalities = len(config["data"]["modalities"])
    return DualEncoder(
        in_channels_per_modality=1,
        num_modalities=num_modalities,
        out_channels=config["model"]["out_channels"],
        features=backbone_config.get("features", [32, 64, 128, 256, 512]),
        norm=backbone_config.get("norm", "instance"),
        fusion_type=fusion_config.get("type", "concat"),
        dropout=config["model"].get("head", {}).get("dropout", 0.0),
    )
