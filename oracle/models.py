"""Oracle (CPU, torch.nn.functional) restatement of the reference's model forward passes.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Every function takes a
``state_dict`` keyed with the reference's parameter names
and an input in the reference's NCDHW layout, and evaluates the arithmetic of
the cited reference code with plain functional ops.  ``dtype`` may be
torch.float32 (default, what the reference computes in) or torch.float64.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _p(sd: SD, key: str, dtype) -> Tensor:
    return sd[key].detach().to("cpu", dtype)


def _act(x: Tensor, activation: str) -> Tensor:
    # reference: src/models/backbones/unet.py:43-51
    if activation == "leaky_relu":
        return F.leaky_relu(x, 0.2)
    if activation == "gelu":
        return F.gelu(x)
    return F.relu(x)


def _norm3d(sd: SD, key: str, x: Tensor, norm: str) -> Tensor:
    """norm1 / norm2 of ConvBlock3D — unet.py:29-41: InstanceNorm3d(affine=False) | BatchNorm3d (eval: running statistics)
    | GroupNorm(8, C) | Identity.  All with eps = 1e-5 (PyTorch defaults)."""
    dt = x.dtype
    if norm == "instance":
        return F.instance_norm(x, eps=1e-5)
    if norm == "group":
        return F.group_norm(x, 8, _p(sd, key + ".weight", dt), _p(sd, key + ".bias", dt), eps=1e-5)
    if norm == "batch":
        return F.batch_norm(x, _p(sd, key + ".running_mean", dt), _p(sd, key + ".running_var", dt),
                            _p(sd, key + ".weight", dt), _p(sd, key + ".bias", dt), training=False, eps=1e-5)
    return x


def conv_block3d(sd: SD, prefix: str, x: Tensor, activation: str = "relu", norm: str = "instance") -> Tensor:
    """ConvBlock3D.forward — src/models/backbones/unet.py:53-60.

    Conv3d(k3,p1,bias) -> norm (InstanceNorm3d(affine=False, eps=1e-5) by default) -> act, twice.
    """
    dt = x.dtype
    for i in (1, 2):
        x = F.conv3d(x, _p(sd, f"{prefix}.conv{i}.weight", dt), _p(sd, f"{prefix}.conv{i}.bias", dt), padding=1)
        x = _norm3d(sd, f"{prefix}.norm{i}", x, norm)
        x = _act(x, activation)
    return x


def down_block3d(sd: SD, prefix: str, x: Tensor, norm: str = "instance") -> Tensor:
    """DownBlock3D.forward — unet.py:76-79 (MaxPool3d(2) then ConvBlock3D)."""
    return conv_block3d(sd, f"{prefix}.conv", F.max_pool3d(x, 2), norm=norm)


def up_block3d(sd: SD, prefix: str, x: Tensor, skip: Tensor, norm: str = "instance") -> Tensor:
    """UpBlock3D.forward — unet.py:104-113 (ConvTranspose3d k2 s2, cat([up, skip]), ConvBlock3D)."""
    dt = x.dtype
    x = F.conv_transpose3d(x, _p(sd, f"{prefix}.up.weight", dt), _p(sd, f"{prefix}.up.bias", dt), stride=2)
    if x.shape != skip.shape:
        x = F.interpolate(x, size=skip.shape[2:], mode="trilinear", align_corners=True)
    x = torch.cat([x, skip], dim=1)
    return conv_block3d(sd, f"{prefix}.conv", x, norm=norm)


def _num_levels(sd: SD, pattern: str) -> int:
    n = 0
    while pattern.format(n) in sd:
        n += 1
    return n


def _unet_norm_kind(sd: SD, prefix: str) -> str:
    """model.backbone.norm as it shows in the state_dict: BatchNorm3d has running statistics, GroupNorm only an affine
    pair, InstanceNorm3d(affine=False) / Identity nothing (those two are told apart by the caller: default instance)."""
    if prefix + "init_conv.norm1.running_mean" in sd:
        return "batch"
    if prefix + "init_conv.norm1.weight" in sd:
        return "group"
    return "instance"


def unet3d_forward(sd: SD, x: Tensor, prefix: str = "backbone.", dtype=torch.float32,
                   return_features: bool = False, norm: str = None):
    """UNet3D.forward — src/models/backbones/unet.py:165-200 (dropout = identity / eval).  norm: model.backbone.norm
    ("instance" | "batch" | "group" | anything else = Identity); None = read it off the state_dict."""
    x = x.detach().to("cpu", dtype)
    if norm is None:
        norm = _unet_norm_kind(sd, prefix)
    n_enc = _num_levels(sd, prefix + "encoders.{}.conv.conv1.weight")
    x = conv_block3d(sd, prefix + "init_conv", x, norm=norm)
    feats = [x]
    for i in range(n_enc):
        x = down_block3d(sd, f"{prefix}encoders.{i}", x, norm=norm)
        feats.append(x)
    feats = feats[:-1]
    for j, skip in enumerate(reversed(feats)):
        x = up_block3d(sd, f"{prefix}decoders.{j}", x, skip, norm=norm)
    x = F.conv3d(x, _p(sd, prefix + "out_conv.weight", dtype), _p(sd, prefix + "out_conv.bias", dtype))
    if return_features:
        return x, feats
    return x


def cross_modal_attention(sd: SD, prefix: str, feats: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """CrossModalAttention.forward — src/models/backbones/dual_encoder.py:243-254.

    Returns (fused, weights[B, M]).  Pooled vector is modality-major (m*C + c).
    """
    dt = feats[0].dtype
    pooled = torch.cat([f.mean(dim=(2, 3, 4)) for f in feats], dim=1)  # [B, M*C]
    h = F.relu(F.linear(pooled, _p(sd, f"{prefix}.attention.2.weight", dt), _p(sd, f"{prefix}.attention.2.bias", dt)))
    w = torch.softmax(F.linear(h, _p(sd, f"{prefix}.attention.4.weight", dt), _p(sd, f"{prefix}.attention.4.bias", dt)), dim=1)
    fused = sum(w[:, m].view(-1, 1, 1, 1, 1) * f for m, f in enumerate(feats))
    return fused, w


def dual_encoder_forward(sd: SD, x: Tensor, fusion_type: str, prefix: str = "backbone.",
                         dtype=torch.float32, return_features: bool = False):
    """DualEncoder.forward — src/models/backbones/dual_encoder.py:112-199."""
    x = x.detach().to("cpu", dtype)
    M = x.shape[1]
    n_blk = _num_levels(sd, prefix + "encoders.0.blocks.{}.conv.conv1.weight")
    all_feats: List[List[Tensor]] = []
    for m in range(M):
        f = conv_block3d(sd, f"{prefix}encoders.{m}.init_conv", x[:, m:m + 1])
        fl = [f]
        for i in range(n_blk):
            f = down_block3d(sd, f"{prefix}encoders.{m}.blocks.{i}", f)
            fl.append(f)
        all_feats.append(fl)
    fused = []
    for lvl in range(n_blk + 1):
        lf = [all_feats[m][lvl] for m in range(M)]
        if fusion_type == "concat":  # dual_encoder.py:179-182
            ff = F.conv3d(torch.cat(lf, 1), _p(sd, f"{prefix}fusion_proj.{lvl}.weight", dtype),
                          _p(sd, f"{prefix}fusion_proj.{lvl}.bias", dtype))
        elif fusion_type == "add":  # :184-186
            ff = sum(lf)
        elif fusion_type == "attention":  # :188-191
            ff, _ = cross_modal_attention(sd, f"{prefix}fusion_layers.{lvl}", lf)
        else:  # :193-195 — 'early', 'late', 'cross_attention', ... all mean
            ff = torch.stack(lf).mean(dim=0)
        fused.append(ff)
    y = fused[-1]
    for j, skip in enumerate(reversed(fused[:-1])):
        y = up_block3d(sd, f"{prefix}decoder.{j}", y, skip)
    y = F.conv3d(y, _p(sd, prefix + "out_conv.weight", dtype), _p(sd, prefix + "out_conv.bias", dtype))
    if return_features:
        return y, {"encoder_features": all_feats, "fused_features": fused}
    return y


def cross_attention_fusion(sd: SD, q_feat: Tensor, kv_feat: Tensor, num_heads: int = 4,
                           prefix: str = "", dtype=torch.float32) -> Tensor:
    """CrossAttentionFusion.forward — src/models/fusion/attention_fusion.py:120-164."""
    q_feat = q_feat.detach().to("cpu", dtype)
    kv_feat = kv_feat.detach().to("cpu", dtype)
    B, C = q_feat.shape[:2]
    hd = C // num_heads
    proj = lambda name, t: F.conv3d(t, _p(sd, f"{prefix}{name}.weight", dtype), _p(sd, f"{prefix}{name}.bias", dtype))
    Q = proj("q_proj", q_feat).reshape(B, num_heads, hd, -1)
    K = proj("k_proj", kv_feat).reshape(B, num_heads, hd, -1)
    V = proj("v_proj", kv_feat).reshape(B, num_heads, hd, -1)
    attn = torch.einsum("bhdn,bhdm->bhnm", Q, K) * (hd ** -0.5)
    attn = torch.softmax(attn, dim=-1)
    out = torch.einsum("bhnm,bhdm->bhdn", attn, V).reshape(q_feat.shape)
    out = proj("out_proj", out)
    return F.instance_norm(q_feat + out, eps=1e-5)


def bidirectional_cross_attention(sd: SD, f1: Tensor, f2: Tensor, num_heads: int = 4, prefix: str = "",
                                  dtype=torch.float32) -> Tensor:
    """BidirectionalCrossAttention.forward — attention_fusion.py:193-216 (both directions, cat, 1x1 conv, IN, ReLU)."""
    a = cross_attention_fusion(sd, f1, f2, num_heads, prefix + "cross_attn_1to2.", dtype)
    b = cross_attention_fusion(sd, f2, f1, num_heads, prefix + "cross_attn_2to1.", dtype)
    y = F.conv3d(torch.cat([a, b], 1), _p(sd, prefix + "fusion.0.weight", dtype), _p(sd, prefix + "fusion.0.bias", dtype))
    return F.relu(F.instance_norm(y, eps=1e-5))


def suv_guided_attention(sd: SD, ct_features: Tensor, pet_suv: Tensor, prefix: str = "", dtype=torch.float32) -> Tensor:
    """SUVGuidedAttention.forward — fusion/attention_fusion.py:266-295: resize PET to the CT feature grid, soft SUV mask
    sigmoid((suv - threshold) * 2), spatial attention Conv3d(1,16,3)+ReLU+Conv3d(16,1,3)+Sigmoid, CT * (1 + attention),
    Conv3d(C,C,1) + InstanceNorm3d."""
    ct = ct_features.detach().to("cpu", dtype)
    pet = pet_suv.detach().to("cpu", dtype)
    if pet.shape[2:] != ct.shape[2:]:
        pet = F.interpolate(pet, size=ct.shape[2:], mode="trilinear", align_corners=True)
    mask = torch.sigmoid((pet - _p(sd, prefix + "threshold", dtype)) * 2)
    a = F.relu(F.conv3d(mask, _p(sd, prefix + "spatial_attn.0.weight", dtype), _p(sd, prefix + "spatial_attn.0.bias", dtype), padding=1))
    a = torch.sigmoid(F.conv3d(a, _p(sd, prefix + "spatial_attn.2.weight", dtype), _p(sd, prefix + "spatial_attn.2.bias", dtype), padding=1))
    y = ct * (1 + a)
    y = F.conv3d(y, _p(sd, prefix + "feature_mod.0.weight", dtype), _p(sd, prefix + "feature_mod.0.bias", dtype))
    return F.instance_norm(y, eps=1e-5)


def attention_fusion(sd: SD, feats: Sequence[Tensor], prefix: str = "", dtype=torch.float32) -> Tensor:
    """AttentionFusion.forward — attention_fusion.py:48-74 (same maths as CrossModalAttention)."""
    feats = [f.detach().to("cpu", dtype) for f in feats]
    pooled = torch.cat([f.mean(dim=(2, 3, 4)) for f in feats], dim=1)
    h = F.relu(F.linear(pooled, _p(sd, f"{prefix}fc.0.weight", dtype), _p(sd, f"{prefix}fc.0.bias", dtype)))
    w = torch.softmax(F.linear(h, _p(sd, f"{prefix}fc.2.weight", dtype), _p(sd, f"{prefix}fc.2.bias", dtype)), dim=1)
    return sum(w[:, m].view(-1, 1, 1, 1, 1) * f for m, f in enumerate(feats))


def model_forward(sd: SD, x: Tensor, model_name: str = "unet", fusion_type: str = "early",
                  dtype=torch.float32) -> Tensor:
    """MultiModalSegmentationModel.forward — src/models/build.py:49-64 (dispatch on MODEL_REGISTRY name)."""
    if model_name in ("unet", "unet3d"):
        return unet3d_forward(sd, x, dtype=dtype)
    if model_name == "dual_encoder":
        return dual_encoder_forward(sd, x, fusion_type, dtype=dtype)
    raise ValueError(f"oracle has no restatement for model {model_name!r}")
