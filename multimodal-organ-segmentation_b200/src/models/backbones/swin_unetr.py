"""Swin UNETR backbone — drop-in for the reference's src/models/backbones/swin_unetr.py.

The reference class is a wrapper whose ctor builds `monai.networks.nets.SwinUNETR` into `self.model` (swin_unetr.py:80-96)
and whose forward is `self.model(x)` (:117).  MONAI is not a dependency here: `self.model` is `SwinUNETRNet`, a parameter
container with MONAI 1.3's attribute tree and state_dict keys (`model.swinViT.layers1.0.blocks.0.attn.qkv.weight`,
`model.encoder1.layer.conv1.conv.weight`, `model.decoder5.transp_conv.conv.weight`, `model.out.conv.conv.bias`, ...), so
a checkpoint written by the reference loads unchanged.  The arithmetic runs in the sm_100a kernels (swin_engine.py);
there is no PyTorch / CPU route.
"""
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from ....swin_engine import SWIN_MODES, SwinUNETREngine
from ....swin_train import swin_unetr_train_forward


class _ConvOnly(nn.Module):
    """monai.networks.blocks.Convolution(conv_only=True): the bare conv lives under `.conv`."""

    def __init__(self, conv: nn.Module):
        super().__init__()
        self.conv = conv


class UnetResBlock(nn.Module):
    """Parameters of MONAI's UnetResBlock (3x3x3 convs without bias; InstanceNorm3d(affine=False) holds no state)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = _ConvOnly(nn.Conv3d(in_channels, out_channels, 3, padding=1, bias=False))
        self.conv2 = _ConvOnly(nn.Conv3d(out_channels, out_channels, 3, padding=1, bias=False))
        if in_channels != out_channels:   # MONAI only creates conv3 / norm3 on a channel change
            self.conv3 = _ConvOnly(nn.Conv3d(in_channels, out_channels, 1, bias=False))


class UnetrBasicBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.layer = UnetResBlock(in_channels, out_channels)


class UnetrUpBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.transp_conv = _ConvOnly(nn.ConvTranspose3d(in_channels, out_channels, 2, stride=2, bias=False))
        self.conv_block = UnetResBlock(2 * out_channels, out_channels)


class UnetOutBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = _ConvOnly(nn.Conv3d(in_channels, out_channels, 1, bias=True))


def _relative_position_index(ws: Sequence[int]) -> torch.Tensor:
    coords = torch.stack(torch.meshgrid(*[torch.arange(w) for w in ws], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    for i in range(3):
        rel[:, :, i] += ws[i] - 1
    rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
    rel[:, :, 1] *= 2 * ws[2] - 1
    return rel.sum(-1)


class WindowAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, window_size: Sequence[int], qkv_bias: bool = True):
        super().__init__()
        self.dim, self.num_heads, self.window_size = dim, num_heads, tuple(window_size)
        n_rel = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_rel, num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        # kept for checkpoint compatibility; the kernel derives the same index arithmetically (swin.cu)
        self.register_buffer("relative_position_index", _relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class MLPBlock(nn.Module):
    def __init__(self, dim: int, mlp_dim: int):
        super().__init__()
        self.linear1 = nn.Linear(dim, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, dim)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim: int, num_heads: int, window_size: Sequence[int], mlp_ratio: float = 4.0, qkv_bias: bool = True):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, num_heads, window_size, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = MLPBlock(dim, int(dim * mlp_ratio))


class PatchMerging(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.reduction = nn.Linear(8 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(8 * dim)


class BasicLayer(nn.Module):
    def __init__(self, dim: int, depth: int, num_heads: int, window_size: Sequence[int]):
        super().__init__()
        self.blocks = nn.ModuleList([SwinTransformerBlock(dim, num_heads, window_size) for _ in range(depth)])
        self.downsample = PatchMerging(dim)


class PatchEmbed(nn.Module):
    def __init__(self, in_chans: int, embed_dim: int):
        super().__init__()
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=2, stride=2)


class SwinTransformer(nn.Module):
    def __init__(self, in_chans: int, embed_dim: int, window_size: Sequence[int], depths: Sequence[int],
                 num_heads: Sequence[int]):
        super().__init__()
        self.patch_embed = PatchEmbed(in_chans, embed_dim)
        for i in range(4):
            setattr(self, f"layers{i + 1}",
                    nn.ModuleList([BasicLayer(embed_dim << i, depths[i], num_heads[i], window_size)]))


class SwinUNETRNet(nn.Module):
    """monai.networks.nets.SwinUNETR's parameter tree (3-D, v1, downsample="merging", res_block=True)."""

    def __init__(self, in_channels: int, out_channels: int, feature_size: int = 48, depths: Sequence[int] = (2, 2, 2, 2),
                 num_heads: Sequence[int] = (3, 6, 12, 24), normalize: bool = True):
        super().__init__()
        if feature_size % 12:
            raise ValueError("feature_size should be divisible by 12.")   # MONAI's own check
        if len(depths) != 4 or len(num_heads) != 4:
            raise ValueError("SwinUNETR has four stages: depths and num_heads need four entries")
        for i in range(4):
            if (feature_size << i) != 16 * num_heads[i]:
                raise NotImplementedError(
                    f"stage {i + 1}: dim {feature_size << i} / heads {num_heads[i]} is not a head_dim of 16 — the sm_100a "
                    "window-attention kernel is built for MONAI's default geometry (feature_size = 16 * num_heads[0])")
        self.in_channels, self.out_channels, self.feature_size = in_channels, out_channels, feature_size
        self.normalize = normalize
        self.window_size = (7, 7, 7)
        F = feature_size
        self.swinViT = SwinTransformer(in_channels, F, self.window_size, depths, num_heads)
        self.encoder1 = UnetrBasicBlock(in_channels, F)
        self.encoder2 = UnetrBasicBlock(F, F)
        self.encoder3 = UnetrBasicBlock(2 * F, 2 * F)
        self.encoder4 = UnetrBasicBlock(4 * F, 4 * F)
        self.encoder10 = UnetrBasicBlock(16 * F, 16 * F)
        self.decoder5 = UnetrUpBlock(16 * F, 8 * F)
        self.decoder4 = UnetrUpBlock(8 * F, 4 * F)
        self.decoder3 = UnetrUpBlock(4 * F, 2 * F)
        self.decoder2 = UnetrUpBlock(2 * F, F)
        self.decoder1 = UnetrUpBlock(F, F)
        self.out = UnetOutBlock(F, out_channels)

    # what the sliding-window inferer reads to fuse the 1x1x1 head into its blend kernel
    @property
    def out_conv(self) -> nn.Conv3d:
        return self.out.conv.conv

    @property
    def features(self) -> List[int]:
        return [self.feature_size << i for i in range(5)]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise RuntimeError("call the SwinUNETR wrapper: the forward runs in the sm_100a engine")


class SwinUNETR(nn.Module):
    """Same constructor / forward / helper signatures as the reference wrapper (swin_unetr.py:20-176)."""

    def __init__(self, img_size: Tuple[int, int, int] = (96, 96, 96), in_channels: int = 1, out_channels: int = 8,
                 feature_size: int = 48, depths: Sequence[int] = (2, 2, 2, 2), num_heads: Sequence[int] = (3, 6, 12, 24),
                 norm_name: str = "instance", drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 dropout_path_rate: float = 0.0, normalize: bool = True, use_checkpoint: bool = False, spatial_dims: int = 3,
                 downsample: str = "merging", use_v2: bool = False, pretrained: Optional[str] = None, **kwargs):
        super().__init__()
        unsupported = []
        if spatial_dims != 3:
            unsupported.append(f"spatial_dims={spatial_dims}")
        if norm_name != "instance":
            unsupported.append(f"norm_name={norm_name!r}")
        if downsample != "merging":
            unsupported.append(f"downsample={downsample!r}")
        if use_v2:
            unsupported.append("use_v2=True")
        if unsupported:
            raise NotImplementedError("SwinUNETR options without an sm_100a kernel: " + ", ".join(unsupported))
        # dropout / drop-path: identity in eval; training needs rates of 0 (the reference's defaults)
        self.drop_rates = (drop_rate, attn_drop_rate, dropout_path_rate)
        self.img_size = img_size
        self.in_channels, self.out_channels, self.feature_size = in_channels, out_channels, feature_size
        self.model = SwinUNETRNet(in_channels, out_channels, feature_size, depths, num_heads, normalize)
        self.numeric_mode = "fp16"
        if pretrained is not None:
            self.load_pretrained(pretrained)

    def set_numeric_mode(self, mode: str) -> "SwinUNETR":
        """'fp16' (default) | 'bf16' operands; names of the UNet ladder that imply more bits map to fp16."""
        mode = mode if mode in SWIN_MODES else "fp16"
        if mode != self.numeric_mode:
            self.numeric_mode = mode
            self.__dict__.pop("_engine", None)
        return self

    def engine(self) -> SwinUNETREngine:
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = self.__dict__["_engine"] = SwinUNETREngine(self.model, self.numeric_mode)
        return eng

    def forward(self, x: torch.Tensor, return_features: bool = False
                ) -> Union[torch.Tensor, Tuple[torch.Tensor, List[torch.Tensor]]]:
        if not x.is_cuda:
            raise RuntimeError("mmseg_b200 modules run on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training: every op is an autograd.Function over the bf16 kernels (swin_train.py)
            if return_features:
                raise NotImplementedError("return_features is an inference-path option")
            if self.training and any(r > 0 for r in self.drop_rates):
                raise NotImplementedError("SwinUNETR dropout / drop-path in training mode is not built (rates must be 0)")
            return swin_unetr_train_forward(self.model, x)
        out = self.engine().forward(x)
        if return_features:
            n, _, Z, Y, X = x.shape
            return out, self.engine().hidden_states(n, Z, Y, X, x.device)
        return out

    def load_pretrained(self, path: str) -> None:
        state_dict = torch.load(path, map_location="cpu")
        for key in ("model_state_dict", "state_dict"):
            if key in state_dict:
                state_dict = state_dict[key]
                break
        missing, unexpected = self.model.load_state_dict(state_dict, strict=False)
        if missing:
            print(f"Missing keys: {len(missing)}")
        if unexpected:
            print(f"Unexpected keys: {len(unexpected)}")

    def get_encoder(self) -> nn.Module:
        return self.model.swinViT

    def get_decoder(self) -> nn.Module:
        m = self.model
        return nn.ModuleList([m.decoder5, m.decoder4, m.decoder3, m.decoder2, m.decoder1])

    @property
    def encoder_channels(self) -> List[int]:
        return [self.feature_size << i for i in range(5)]


def build_swin_unetr(config: Dict[str, Any]) -> SwinUNETR:
    """Same config keys as the reference's build_swin_unetr (swin_unetr.py:179-200)."""
    bc = config.get("model", {}).get("backbone", {})
    return SwinUNETR(
        img_size=tuple(bc.get("img_size", [96, 96, 96])),
        in_channels=config["model"]["in_channels"],
        out_channels=config["model"]["out_channels"],
        feature_size=bc.get("feature_size", 48),
        depths=tuple(bc.get("depths", [2, 2, 2, 2])),
        num_heads=tuple(bc.get("num_heads", [3, 6, 12, 24])),
        drop_rate=config["model"].get("head", {}).get("dropout", 0.0),
        use_checkpoint=config.get("training", {}).get("use_checkpoint", False),
    )
