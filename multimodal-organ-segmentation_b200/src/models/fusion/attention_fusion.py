"""Attention-based fusion — drop-in mirror of the reference's src/models/fusion/attention_fusion.py
(AttentionFusion, CrossAttentionFusion, BidirectionalCrossAttention; same constructors, parameter names, forward
signatures).  Forward-only in the sm_100a kernels (bf16 operands, fp32 accumulation / softmax):
  * q/k/v/out 1x1 projections  -> tcgen05 implicit-GEMM conv kernel (k and v as ONE conv over the key/value features);
  * softmax(QK^T)V              -> fused flash-style tcgen05 attention kernel (csrc/attention.cu), no N x N matrix;
  * InstanceNorm3d(q + out)     -> add + statistics kernel, finalize, normalise (no activation).
SUVGuidedAttention (never instantiated by the reference, needs a sigmoid gate kernel) is not built.
"""
from typing import List

import torch
import torch.nn as nn

from ..backbones.unet import _require_cuda, _no_autograd
from .... import kernels as K
from .... import _lib
from ....engine import ConvRunner
from ....kernels import Blocked


class AttentionFusion(nn.Module):
    """reference attention_fusion.py:12-74 (SE-style modality gate; same maths as DualEncoder's CrossModalAttention)."""

    def __init__(self, in_channels: int, num_modalities: int = 2, reduction: int = 4):
        super().__init__()
        self.in_channels = in_channels
        self.num_modalities = num_modalities
        self.global_pool = nn.AdaptiveAvgPool3d(1)
        self.fc = nn.Sequential(
            nn.Linear(in_channels * num_modalities, in_channels * num_modalities // reduction),
            nn.ReLU(inplace=True),
            nn.Linear(in_channels * num_modalities // reduction, num_modalities),
            nn.Softmax(dim=1),
        )
        self.out_channels = in_channels

    def forward(self, features: List[torch.Tensor]) -> torch.Tensor:
        _require_cuda(features[0])
        _no_autograd(self, features[0])
        B, C, Z, Y, X = features[0].shape
        M = len(features)
        if C % 16:
            raise NotImplementedError("AttentionFusion kernels need channels % 16 == 0")
        with torch.no_grad():
            st = Blocked(B, M * C, Z, Y, X, False, features[0].device)
            for m, f in enumerate(features):
                K.pack_ncdhw(f.contiguous().float(), st, c0=m * C)
            pooled = K.channel_mean(st, 0, M * C)
            w = K.gate_mlp(pooled, self.fc[0].weight, self.fc[0].bias, self.fc[2].weight, self.fc[2].bias)
            out = Blocked(B, C, Z, Y, X, False, features[0].device)
            K.modality_combine(st, M, C, out, 0, w)
            return out.to_ncdhw()


def _pad_heads_out(w: torch.Tensor, b: torch.Tensor, heads: int, hd: int, hdp: int):
    """[C, Cin] projection whose OUTPUT channels are re-laid as heads x hdp (zero rows for the padding)."""
    C, cin = w.shape[0], w.shape[1]
    wp = torch.zeros((heads * hdp, cin), dtype=torch.float32, device=w.device)
    bp = torch.zeros(heads * hdp, dtype=torch.float32, device=w.device)
    wv = w.detach().float().reshape(heads, hd, cin)
    wp.view(heads, hdp, cin)[:, :hd] = wv
    bp.view(heads, hdp)[:, :hd] = b.detach().float().view(heads, hd)
    return wp, bp


class CrossAttentionFusion(nn.Module):
    """reference attention_fusion.py:77-164."""

    def __init__(self, in_channels: int, num_heads: int = 4, dropout: float = 0.0):
        super().__init__()
        self.in_channels = in_channels
        self.num_heads = num_heads
        self.head_dim = in_channels // num_heads
        assert in_channels % num_heads == 0, "in_channels must be divisible by num_heads"
        self.q_proj = nn.Conv3d(in_channels, in_channels, kernel_size=1)
        self.k_proj = nn.Conv3d(in_channels, in_channels, kernel_size=1)
        self.v_proj = nn.Conv3d(in_channels, in_channels, kernel_size=1)
        self.out_proj = nn.Conv3d(in_channels, in_channels, kernel_size=1)
        self.dropout = nn.Dropout(dropout)
        self.norm = nn.InstanceNorm3d(in_channels)
        self.out_channels = in_channels
        self._runner = None

    def _packed(self):
        C, h, hd = self.in_channels, self.num_heads, self.head_dim
        hdp = max(16, (hd + 15) // 16 * 16)
        if hdp not in (16, 32, 64, 128):
            raise NotImplementedError(f"CrossAttentionFusion: head_dim {hd} has no sm_100a kernel (<= 128, padded to 16/32/64/128)")
        ver = tuple(p._version for p in self.parameters())
        c = self.__dict__.get("_pack_cache")
        if c is not None and c[0] == ver:
            return c[1]
        r2 = lambda conv: conv.weight.detach().float().reshape(C, C)
        wq, bq = _pad_heads_out(r2(self.q_proj), self.q_proj.bias, h, hd, hdp)
        wk, bk = _pad_heads_out(r2(self.k_proj), self.k_proj.bias, h, hd, hdp)
        wv, bv = _pad_heads_out(r2(self.v_proj), self.v_proj.bias, h, hd, hdp)
        wkv, bkv = torch.cat([wk, wv]), torch.cat([bk, bv])
        wo = torch.zeros((C, h * hdp), dtype=torch.float32, device=wq.device)   # input channels in the padded head layout
        wo.view(C, h, hdp)[:, :, :hd] = r2(self.out_proj).view(C, h, hd)
        f5 = lambda w: w.reshape(w.shape[0], w.shape[1], 1, 1, 1)
        packs = {"q": K.pack_conv_weight(f5(wq), bq, False, [C]), "kv": K.pack_conv_weight(f5(wkv), bkv, False, [C]),
                 "o": K.pack_conv_weight(f5(wo), self.out_proj.bias, False, [h * hdp]), "hdp": hdp}
        self.__dict__["_pack_cache"] = (ver, packs)
        return packs

    def forward_blocked(self, q_in: Blocked, kv_in: Blocked, dst: Blocked, dst_c0: int = 0) -> None:
        """q_in / kv_in: blocked bf16 features (C channels at block 0); writes InstanceNorm(q + attention) into dst."""
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("attention dropout (p > 0, train mode) is not built in the fused kernel")
        C, h, hd = self.in_channels, self.num_heads, self.head_dim
        P = self._packed()
        hdp = P["hdp"]
        n, Z, Y, X = q_in.n_img, q_in.Z, q_in.Y, q_in.X
        dev = q_in.t.device
        if self._runner is None:
            self._runner = ConvRunner(False, dev)
        r = self._runner
        qb = Blocked(n, h * hdp, Z, Y, X, False, dev)
        kvb = Blocked(n, 2 * h * hdp, Z, Y, X, False, dev)
        ab = Blocked(n, h * hdp, Z, Y, X, False, dev)
        ob = Blocked(n, C, Z, Y, X, False, dev)
        r.conv_act(q_in, [(0, C)], P["q"], qb)
        r.conv_act(kv_in, [(0, C)], P["kv"], kvb)
        K.cross_attention(qb, 0, kvb, 0, h * hdp, ab, 0, h, hdp, float(hd) ** -0.5)
        r.conv_act(ab, [(0, h * hdp)], P["o"], ob)
        y, partial, n_chunks = K.add_stats(q_in, 0, ob, 0, C)
        mr = torch.empty((n, C, 2), dtype=torch.float32, device=dev)
        K.instnorm_finalize(partial, n, n_chunks, C, Z * Y * X, mr, eps=self.norm.eps)
        K.instnorm_act_apply(y, True, mr, n, C, Z, Y, X, dst, dst_c0, slope=1.0)   # slope 1 = no activation

    def forward(self, query_features: torch.Tensor, key_value_features: torch.Tensor) -> torch.Tensor:
        _require_cuda(query_features)
        _no_autograd(self, query_features)
        B, C, Z, Y, X = query_features.shape
        if C % 16:
            raise NotImplementedError("CrossAttentionFusion kernels need channels % 16 == 0")
        with torch.no_grad():
            dev = query_features.device
            q_in = Blocked(B, C, Z, Y, X, False, dev)
            kv_in = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw(query_features.contiguous().float(), q_in)
            K.pack_ncdhw(key_value_features.contiguous().float(), kv_in)
            dst = Blocked(B, C, Z, Y, X, False, dev)
            self.forward_blocked(q_in, kv_in, dst)
            return dst.to_ncdhw()


class BidirectionalCrossAttention(nn.Module):
    """reference attention_fusion.py:167-216."""

    def __init__(self, in_channels: int, num_heads: int = 4, dropout: float = 0.0):
        super().__init__()
        self.cross_attn_1to2 = CrossAttentionFusion(in_channels, num_heads, dropout)
        self.cross_attn_2to1 = CrossAttentionFusion(in_channels, num_heads, dropout)
        self.fusion = nn.Sequential(
            nn.Conv3d(in_channels * 2, in_channels, kernel_size=1),
            nn.InstanceNorm3d(in_channels),
            nn.ReLU(inplace=True),
        )
        self.out_channels = in_channels
        self._runner = None

    def forward(self, features_1: torch.Tensor, features_2: torch.Tensor) -> torch.Tensor:
        _require_cuda(features_1)
        _no_autograd(self, features_1)
        B, C, Z, Y, X = features_1.shape
        with torch.no_grad():
            dev = features_1.device
            f1 = Blocked(B, C, Z, Y, X, False, dev)
            f2 = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw(features_1.contiguous().float(), f1)
            K.pack_ncdhw(features_2.contiguous().float(), f2)
            cat = Blocked(B, 2 * C, Z, Y, X, False, dev)          # torch.cat([attn_1to2, attn_2to1], 1) for free
            self.cross_attn_1to2.forward_blocked(f1, f2, cat, 0)
            self.cross_attn_2to1.forward_blocked(f2, f1, cat, C)
            if self._runner is None:
                self._runner = ConvRunner(False, dev)
            conv = self.fusion[0]
            pw = K.pack_conv_weight(conv.weight, None, False, [C, C], use_bias=False)   # bias cancelled by the norm
            out = Blocked(B, C, Z, Y, X, False, dev)
            self._runner.conv_norm_act(cat, [(0, C), (C, C)], pw, out)
            return out.to_ncdhw()


class SUVGuidedAttention(nn.Module):
    """reference attention_fusion.py:219-295: high-SUV regions of the PET image gate the CT features.

    Same constructor / parameters / state_dict (threshold buffer or parameter, spatial_attn.{0,2}, feature_mod.0).  The
    arithmetic: trilinear resize (kernel), soft mask folded into the blocked pack, the two 3x3x3 convs on the tcgen05
    kernel (bias + ReLU through the norm-less apply; 1-channel logits as NCDHW fp32), the sigmoid gate CT * (1 + a) folded
    into the pack of the CT features, Conv3d(C, C, 1) + InstanceNorm3d (no activation) through the conv / norm kernels."""

    def __init__(self, in_channels: int, suv_threshold: float = 2.5, learnable_threshold: bool = False):
        super().__init__()
        self.in_channels = in_channels
        if learnable_threshold:
            self.threshold = nn.Parameter(torch.tensor(suv_threshold))
        else:
            self.register_buffer("threshold", torch.tensor(suv_threshold))
        self.spatial_attn = nn.Sequential(
            nn.Conv3d(1, 16, kernel_size=3, padding=1),
            nn.ReLU(inplace=True),
            nn.Conv3d(16, 1, kernel_size=3, padding=1),
            nn.Sigmoid(),
        )
        self.feature_mod = nn.Sequential(
            nn.Conv3d(in_channels, in_channels, kernel_size=1),
            nn.InstanceNorm3d(in_channels),
        )
        self.out_channels = in_channels
        self._runner = None

    def forward(self, ct_features: torch.Tensor, pet_suv: torch.Tensor) -> torch.Tensor:
        _require_cuda(ct_features)
        _no_autograd(self, ct_features)
        if self.in_channels % 16:
            raise NotImplementedError("SUVGuidedAttention kernels need channels % 16 == 0")
        with torch.no_grad():
            dev = ct_features.device
            ct = ct_features.contiguous().float()
            pet = pet_suv.contiguous().float()
            B, C, Z, Y, X = ct.shape
            if tuple(pet.shape[2:]) != (Z, Y, X):
                pet = K.trilinear_resize(pet, (Z, Y, X))
            if self._runner is None:
                self._runner = ConvRunner(False, dev)
            r = self._runner
            # soft SUV mask, packed straight into the first conv's (padded) input
            m = Blocked(B, 16, Z, Y, X, False, dev)
            K.pack_ncdhw_ex(pet, m, pre_sigmoid=(float(self.threshold), 2.0))
            c1, c2 = self.spatial_attn[0], self.spatial_attn[2]
            h = Blocked(B, 16, Z, Y, X, False, dev)
            r.conv_norm_act(m, [(0, 1)], K.pack_conv_weight(c1.weight, c1.bias, False, [1]), h, norm=nn.Identity())
            gate = torch.empty((B, 1, Z, Y, X), dtype=torch.float32, device=dev)       # attention LOGITS (pre-sigmoid)
            pw2 = K.pack_conv_weight(c2.weight, c2.bias, False, [16])
            K.conv3d(h, pw2, K.a_chunk_table(h, [0], [16], False), gate, _lib.OUT_NCDHW_F32)
            # CT * (1 + sigmoid(logits)) folded into the pack of the CT features
            att = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw_ex(ct, att, gate_logits=gate)
            fm = self.feature_mod[0]
            out = Blocked(B, C, Z, Y, X, False, dev)
            # Conv3d(C, C, 1) + InstanceNorm3d without activation: LeakyReLU with slope 1 is the identity
            r.conv_norm_act(att, [(0, C)], K.pack_conv_weight(fm.weight, None, False, [C], use_bias=False), out, slope=1.0)
            return out.to_ncdhw()
