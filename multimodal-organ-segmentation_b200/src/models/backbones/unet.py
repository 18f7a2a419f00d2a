"""3D U-Net backbone — drop-in mirror of the reference's src/models/backbones/unet.py.

Same class names, constructor arguments, attribute tree and state_dict keys/shapes so a
reference checkpoint loads unchanged; the arithmetic runs in the sm_100a kernels (engine.py).  There is no
PyTorch/CPU route: a forward on a non-CUDA tensor, or with an option the kernels do not cover, raises.
"""
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from ....engine import ConvRunner, UNet3DEngine
from ....train_engine import TrainEngine, train_forward
from .... import kernels as K
from .... import _lib
from ....kernels import Blocked
from ....numerics import DEFAULT_INFERENCE_MODE, MODES, mode as numeric_mode

# numeric mode of the INFERENCE path (numerics.py): the fastest rung of the ladder that meets north_star's gates.  The
# training path (TrainEngine) always computes in bf16 (north_star: bf16 training step), whatever this says.
_DEFAULT_MODE = DEFAULT_INFERENCE_MODE


def _require_cuda(x: torch.Tensor) -> None:
    if not x.is_cuda:
        raise RuntimeError("mmseg_b200 modules run on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _wants_grad(module: nn.Module, x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))


def _no_autograd(module: nn.Module, x: torch.Tensor) -> None:
    """Stand-alone building blocks are forward-only; whole models (UNet3D / DualEncoder) train through TrainEngine."""
    if _wants_grad(module, x):
        raise NotImplementedError(
            "this stand-alone block has no backward of its own: wrap the call in torch.no_grad(), or train the whole "
            "UNet3D / DualEncoder (their backward runs in the sm_100a dgrad / wgrad / norm-backward kernels)")


def _train_step_forward(model: nn.Module, kind: str, x: torch.Tensor) -> torch.Tensor:
    """Forward that records the tape for the kernel backward (parameters get gradients; so does the input volume when
    x.requires_grad: the first layers' dgrad then runs as well)."""
    blk = getattr(model, "init_conv", None)
    if blk is not None and getattr(blk, "activation", "relu") == "gelu":
        raise NotImplementedError("training with activation='gelu' is not built (the backward kernels cover ReLU / LeakyReLU)")
    # model.backbone.norm = batch | group | none trains through the same kernels (train_engine._conv_generic_norm_act), for
    # UNet3D and DualEncoder (every fusion type), with or without Dropout3d
    eng = model.__dict__.get("_train_engine")
    if eng is None:
        eng = model.__dict__["_train_engine"] = TrainEngine(model, kind)
    drop = None
    if model.training and isinstance(model.dropout, nn.Dropout3d) and model.dropout.p > 0:
        # same draw as nn.Dropout3d (feature dropout): one Bernoulli(1-p) per (sample, channel), scaled by 1/(1-p)
        p = model.dropout.p
        noise = torch.empty((x.shape[0], model.features[0]), dtype=torch.float32, device=x.device).bernoulli_(1 - p)
        drop = noise / (1 - p)
    return train_forward(eng, x, drop)


class ConvBlock3D(nn.Module):
    """(Conv3d k3 p1 -> InstanceNorm3d -> act) x 2 — reference unet.py:12-60."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, padding: int = 1,
                 norm: str = "instance", activation: str = "relu"):
        super().__init__()
        self.conv1 = nn.Conv3d(in_channels, out_channels, kernel_size, padding=padding)
        self.conv2 = nn.Conv3d(out_channels, out_channels, kernel_size, padding=padding)
        if norm == "batch":
            self.norm1, self.norm2 = nn.BatchNorm3d(out_channels), nn.BatchNorm3d(out_channels)
        elif norm == "instance":
            self.norm1, self.norm2 = nn.InstanceNorm3d(out_channels), nn.InstanceNorm3d(out_channels)
        elif norm == "group":
            self.norm1, self.norm2 = nn.GroupNorm(8, out_channels), nn.GroupNorm(8, out_channels)
        else:
            self.norm1, self.norm2 = nn.Identity(), nn.Identity()
        if activation == "leaky_relu":
            self.act = nn.LeakyReLU(0.2, inplace=True)
        elif activation == "gelu":
            self.act = nn.GELU()
        else:
            self.act = nn.ReLU(inplace=True)
        self.norm_type = norm
        self.activation = activation if activation in ("leaky_relu", "gelu") else "relu"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.numeric_mode = _DEFAULT_MODE
        self._runner: Optional[ConvRunner] = None
        self._packed = None

    def kernel_supported(self) -> None:
        if self.conv1.kernel_size != (3, 3, 3) or self.conv1.padding != (1, 1, 1) or self.out_channels % 16:
            raise NotImplementedError(
                f"ConvBlock3D(norm={self.norm_type!r}, activation={self.activation!r}, k={self.conv1.kernel_size}, "
                f"C_out={self.out_channels}) has no sm_100a kernel (covered: instance / group / batch(eval) / no norm, "
                "relu / leaky_relu / gelu, k=3 p=1, C_out % 16 == 0)")

    @property
    def slope(self) -> float:
        return 0.2 if self.activation == "leaky_relu" else 0.0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x)
        _no_autograd(self, x)
        self.kernel_supported()
        split = numeric_mode(self.numeric_mode)
        with torch.no_grad():
            x = x.contiguous().float()
            n, c, Z, Y, X = x.shape
            inst = self.norm_type == "instance"   # only InstanceNorm cancels the conv bias
            ver = (self.conv1.weight._version, self.conv2.weight._version, self.conv1.bias._version,
                   self.conv2.bias._version, split)
            if self._packed is None or self._packed[0] != ver:
                self._packed = (ver, K.pack_conv_weight(self.conv1.weight, None if inst else self.conv1.bias, split, [c],
                                                        use_bias=not inst),
                                K.pack_conv_weight(self.conv2.weight, None if inst else self.conv2.bias, split, None,
                                                   use_bias=not inst))
            if self._runner is None or self._runner.split != split:
                self._runner = ConvRunner(split, x.device)
            src = Blocked(n, (c + 15) // 16 * 16, Z, Y, X, split, x.device)
            K.pack_ncdhw(x, src)
            mid = Blocked(n, self.out_channels, Z, Y, X, split, x.device)
            out = Blocked(n, self.out_channels, Z, Y, X, split, x.device)
            gelu = self.activation == "gelu"
            self._runner.conv_norm_act(src, [(0, c)], self._packed[1], mid, slope=self.slope, norm=self.norm1, gelu=gelu)
            self._runner.conv_norm_act(mid, [(0, self.out_channels)], self._packed[2], out, slope=self.slope,
                                       norm=self.norm2, gelu=gelu)
            return out.to_ncdhw()


class DownBlock3D(nn.Module):
    """MaxPool3d(2) -> ConvBlock3D; returns (x_conv, x_pool) — reference unet.py:63-79."""

    def __init__(self, in_channels: int, out_channels: int, norm: str = "instance"):
        super().__init__()
        self.pool = nn.MaxPool3d(2)
        self.conv = ConvBlock3D(in_channels, out_channels, norm=norm)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _require_cuda(x)
        _no_autograd(self, x)
        # standalone use only (UNet3D / DualEncoder run the fused engine): pooling is a layout-only op here
        x_pool = torch.nn.functional.max_pool3d(x, 2)
        return self.conv(x_pool), x_pool


class UpBlock3D(nn.Module):
    """ConvTranspose3d(k2,s2) -> cat([up, skip]) -> ConvBlock3D — reference unet.py:82-113."""

    def __init__(self, in_channels: int, out_channels: int, norm: str = "instance", mode: str = "transpose"):
        super().__init__()
        if mode == "transpose":
            self.up = nn.ConvTranspose3d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        else:
            self.up = nn.Sequential(
                nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True),
                nn.Conv3d(in_channels, in_channels // 2, kernel_size=1),
            )
        self.mode = mode
        self.conv = ConvBlock3D(in_channels, out_channels, norm=norm)
        self.numeric_mode = _DEFAULT_MODE
        self._runner: Optional[ConvRunner] = None
        self._packed = None

    def forward(self, x: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        _require_cuda(x)
        _no_autograd(self, x)
        if self.mode != "transpose":
            raise NotImplementedError("UpBlock3D(mode != 'transpose') is never selected by the reference's builders "
                                      "(unet.py:149) and has no sm_100a kernel")
        self.conv.kernel_supported()
        split = numeric_mode(self.numeric_mode)
        with torch.no_grad():
            x = x.contiguous().float()
            skip = skip.contiguous().float()
            n, c, Z, Y, X = x.shape
            resize = tuple(skip.shape[2:]) != (2 * Z, 2 * Y, 2 * X)     # reference unet.py:108-109
            half = c // 2
            inst = self.conv.norm_type == "instance"   # only InstanceNorm cancels the conv bias
            ver = (self.up.weight._version, self.conv.conv1.weight._version, self.conv.conv2.weight._version,
                   self.conv.conv1.bias._version, self.conv.conv2.bias._version, split)
            if self._packed is None or self._packed[0] != ver:
                self._packed = (ver,
                                K.pack_conv_weight(self.up.weight, self.up.bias, split, None, transposed=True),
                                K.pack_conv_weight(self.conv.conv1.weight, None if inst else self.conv.conv1.bias, split,
                                                   [half, skip.shape[1]], use_bias=not inst),
                                K.pack_conv_weight(self.conv.conv2.weight, None if inst else self.conv.conv2.bias, split,
                                                   None, use_bias=not inst))
            if self._runner is None or self._runner.split != split:
                self._runner = ConvRunner(split, x.device)
            r = self._runner
            src = Blocked(n, c, Z, Y, X, split, x.device)
            K.pack_ncdhw(x, src)
            SZ, SY, SX = (int(v) for v in skip.shape[2:])
            cat = Blocked(n, half + skip.shape[1], SZ, SY, SX, split, x.device)
            K.pack_ncdhw(skip, cat, c0=half)
            if resize:   # ConvTranspose at 2x, then F.interpolate(..., mode="trilinear", align_corners=True) to the skip's size
                up = Blocked(n, half, 2 * Z, 2 * Y, 2 * X, split, x.device)
                r.conv_transpose(src, [(0, c)], self._packed[1], up, 0)
                K.pack_ncdhw(K.trilinear_resize(up.to_ncdhw(), (SZ, SY, SX)), cat, 0)
            else:
                r.conv_transpose(src, [(0, c)], self._packed[1], cat, 0)
            co = self.conv.out_channels
            mid = Blocked(n, co, SZ, SY, SX, split, x.device)
            out = Blocked(n, co, SZ, SY, SX, split, x.device)
            r.conv_norm_act(cat, [(0, half), (half, skip.shape[1])], self._packed[2], mid, slope=self.conv.slope,
                            norm=self.conv.norm1)
            r.conv_norm_act(mid, [(0, co)], self._packed[3], out, slope=self.conv.slope, norm=self.conv.norm2)
            return out.to_ncdhw()


class UNet3D(nn.Module):
    """3D UNet — reference unet.py:116-205 (same ctor / forward / encoder_channels)."""

    def __init__(self, in_channels: int = 1, out_channels: int = 8, features: List[int] = [32, 64, 128, 256, 512],
                 norm: str = "instance", dropout: float = 0.0, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.features = list(features)
        self.init_conv = ConvBlock3D(in_channels, features[0], norm=norm)
        self.encoders = nn.ModuleList()
        for i in range(len(features) - 1):
            self.encoders.append(DownBlock3D(features[i], features[i + 1], norm=norm))
        self.decoders = nn.ModuleList()
        for i in range(len(features) - 1, 0, -1):
            self.decoders.append(UpBlock3D(features[i], features[i - 1], norm=norm))
        self.dropout = nn.Dropout3d(dropout) if dropout > 0 else nn.Identity()
        self.out_conv = nn.Conv3d(features[0], out_channels, kernel_size=1)
        self.numeric_mode = _DEFAULT_MODE
        self._engines: Dict[str, UNet3DEngine] = {}

    def set_numeric_mode(self, mode: str) -> "UNet3D":
        """Inference numeric mode, a rung of the ladder in numerics.py: 'fp16m' (default: fastest mode that meets the
        stated logit / label tolerances), 'bf16' / 'fp16' (single pass, fastest), 'parity' (3-pass split-bf16, ~5e-5),
        'fp16w2', 'fp16a2', 'fp16x3'.  Training always runs the bf16 kernels."""
        self.numeric_mode = numeric_mode(mode).name
        return self

    def engine(self) -> UNet3DEngine:
        e = self._engines.get(self.numeric_mode)
        if e is None:
            e = self._engines[self.numeric_mode] = UNet3DEngine(self, self.numeric_mode)
        return e

    def forward(self, x: torch.Tensor, return_features: bool = False
                ) -> Union[torch.Tensor, Tuple[torch.Tensor, List[torch.Tensor]]]:
        _require_cuda(x)
        self.init_conv.kernel_supported()
        if _wants_grad(self, x):
            if return_features:
                raise NotImplementedError("return_features is an inference-path option")
            return _train_step_forward(self, "unet", x)
        if self.training and isinstance(self.dropout, nn.Dropout3d) and self.dropout.p > 0:
            raise NotImplementedError("train-mode Dropout3d under no_grad: call model.eval() for inference")
        eng = self.engine()
        logits = eng.forward(x)
        if return_features:
            n, _, Z, Y, X = x.shape
            feats = [eng.feature_ncdhw(n, Z, Y, X, l) for l in range(len(self.features) - 1)]
            return logits, feats
        return logits

    @property
    def encoder_channels(self) -> List[int]:
        return self.features


def build_unet3d(config: Dict[str, Any]) -> UNet3D:
    """reference unet.py:208-226."""
    backbone_config = config.get("model", {}).get("backbone", {})
    return UNet3D(
        in_channels=config["model"]["in_channels"],
        out_channels=config["model"]["out_channels"],
        features=backbone_config.get("features", [32, 64, 128, 256, 512]),
        norm=backbone_config.get("norm", "instance"),
        dropout=config["model"].get("head", {}).get("dropout", 0.0),
    )
