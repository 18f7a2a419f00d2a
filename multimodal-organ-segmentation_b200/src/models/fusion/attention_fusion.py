"""Attention-based fusion — drop-in mirror of the reference's src/models/fusion/attention_fusion.py
(AttentionFusion, CrossAttentionFusion, BidirectionalCrossAttention; same constructors, parameter names, forward
signatures), in the sm_100a kernels (bf16 operands, fp32 accumulation / softmax):
  * q/k/v/out 1x1 projections  -> tcgen05 implicit-GEMM conv kernel (k and v as ONE conv over the key/value features);
  * softmax(QK^T)V              -> fused flash-style tcgen05 attention kernel (csrc/attention.cu), no N x N matrix;
  * InstanceNorm3d(q + out)     -> add + statistics kernel, finalize, normalise (no activation).
CrossAttentionFusion also TRAINS through the kernels (_CrossAttentionFunction: norm backward, projection dgrad / wgrad,
flash-style attention backward by recomputation); the other modules of this file are forward-only.
"""
from typing import List, Optional

import torch
import torch.nn as nn

from ..backbones.unet import _require_cuda, _no_autograd, _wants_grad
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

    def forward_blocked(self, q_in: Blocked, kv_in: Blocked, dst: Blocked, dst_c0: int = 0, save: Optional[dict] = None) -> None:
        """q_in / kv_in: blocked bf16 features (C channels at block 0); writes InstanceNorm(q + attention) into dst.
        save (training): receives what backward_blocked needs (projections, attention output, log-sum-exp rows, the
        pre-norm sum as bf16 and its statistics)."""
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
        lse = torch.empty((n, h, Z * Y * X), dtype=torch.float32, device=dev) if save is not None else None
        r.conv_act(q_in, [(0, C)], P["q"], qb)
        r.conv_act(kv_in, [(0, C)], P["kv"], kvb)
        K.cross_attention(qb, 0, kvb, 0, h * hdp, ab, 0, h, hdp, float(hd) ** -0.5, lse=lse)
        r.conv_act(ab, [(0, h * hdp)], P["o"], ob)
        y, partial, n_chunks = K.add_stats(q_in, 0, ob, 0, C)
        mr = torch.empty((n, C, 2), dtype=torch.float32, device=dev)
        K.instnorm_finalize(partial, n, n_chunks, C, Z * Y * X, mr, eps=self.norm.eps)
        K.instnorm_act_apply(y, True, mr, n, C, Z, Y, X, dst, dst_c0, slope=1.0)   # slope 1 = no activation
        if save is not None:
            # the norm backward reads the pre-norm tensor in bf16: an identity pass (mean 0, rstd 1, slope 1) converts it
            ybf = Blocked(n, C, Z, Y, X, False, dev)
            ident = torch.zeros((n, C, 2), dtype=torch.float32, device=dev)
            ident[:, :, 1] = 1.0
            K.instnorm_act_apply(y, True, ident, n, C, Z, Y, X, ybf, 0, slope=1.0)
            save.update(q_in=q_in, kv_in=kv_in, qb=qb, kvb=kvb, ab=ab, lse=lse, ybf=ybf, mr=mr, hdp=hdp)

    def backward_blocked(self, saved: dict, g_out: Blocked, g_c0: int = 0, dq_dst: Optional[Blocked] = None, dq_c0: int = 0,
                         dkv_dst: Optional[Blocked] = None, dkv_c0: int = 0):
        """Backward of forward_blocked in the sm_100a kernels (autograd through attention_fusion.py:138-162 in the
        reference): InstanceNorm backward, out_proj dgrad / wgrad, the flash-style attention backward, q / k / v projection
        dgrad / wgrad, and the residual.  g_out: gradient w.r.t. the module output (blocked bf16, C channels from channel
        g_c0).  Returns (d_query_features, d_key_value_features) as Blocked — written into dq_dst / dkv_dst at the given
        channel offsets when supplied — and {parameter: fp32 gradient}."""
        from ....train_engine import _wrap
        C, h, hd = self.in_channels, self.num_heads, self.head_dim
        hdp = saved["hdp"]
        q_in, kv_in, qb, kvb, ab = saved["q_in"], saved["kv_in"], saved["qb"], saved["kvb"], saved["ab"]
        n, Z, Y, X = q_in.n_img, q_in.Z, q_in.Y, q_in.X
        dev = q_in.t.device
        nvox = Z * Y * X
        HP = h * hdp
        chan_sums = lambda b, c0, ch: (K.channel_mean(b, c0, ch) * float(nvox)).sum(0)
        # d(q + out_proj(attn)) through the InstanceNorm (no activation: slope 1); lands in channels [0, C) of `stack`, whose
        # channels [C, 2C) receive the q_proj input gradient: d_query = their sum (the residual)
        stack = Blocked(n, 2 * C, Z, Y, X, False, dev)
        K.instnorm_act_bwd(saved["ybf"].t, saved["mr"], n, C, Z, Y, X, g_out, g_c0, 1.0, None, 0, stack.t, slope=1.0,
                           dx_cbt=stack.cbt, dx_cb_off=0)
        grads = {}
        W = self._bwd_weights()
        # out_proj
        dwo = K.conv3d_wgrad(ab, [(0, HP)], stack.t, stack.cbt, 0, C, 1, (C, HP, 1, 1, 1))
        grads[self.out_proj.weight] = dwo.view(C, h, hdp)[:, :, :hd].reshape(C, C, 1, 1, 1)
        grads[self.out_proj.bias] = chan_sums(stack, 0, C)
        d_ab = Blocked(n, HP, Z, Y, X, False, dev)
        K.conv3d(stack, W["o"], K.a_chunk_table(stack, [0], [C], False), d_ab.t, _lib.OUT_BLOCKED_BF16, dst_cbt=d_ab.cbt)
        # attention core
        dqb = Blocked(n, HP, Z, Y, X, False, dev)
        dkvb = Blocked(n, 2 * HP, Z, Y, X, False, dev)
        K.cross_attention_bwd(qb, 0, kvb, 0, HP, ab, 0, d_ab, 0, saved["lse"], dqb, 0, dkvb, 0, HP, h, hdp, float(hd) ** -0.5)
        # q projection
        dwq = K.conv3d_wgrad(q_in, [(0, C)], dqb.t, dqb.cbt, 0, HP, 1, (HP, C, 1, 1, 1))
        grads[self.q_proj.weight] = dwq.view(h, hdp, C)[:, :hd].reshape(C, C, 1, 1, 1)
        grads[self.q_proj.bias] = chan_sums(dqb, 0, HP).view(h, hdp)[:, :hd].reshape(C)
        K.conv3d(dqb, W["q"], K.a_chunk_table(dqb, [0], [HP], False), stack.t, _lib.OUT_BLOCKED_BF16, dst_cbt=stack.cbt,
                 dst_cb_off=C // 8)
        d_q = dq_dst if dq_dst is not None else Blocked(n, C, Z, Y, X, False, dev)
        K.modality_combine(stack, 2, C, d_q, dq_c0, None, 1.0)
        # k / v projections (one stacked conv, like the forward)
        dwkv = K.conv3d_wgrad(kv_in, [(0, C)], dkvb.t, dkvb.cbt, 0, 2 * HP, 1, (2 * HP, C, 1, 1, 1))
        bkv = chan_sums(dkvb, 0, 2 * HP)
        grads[self.k_proj.weight] = dwkv[:HP].view(h, hdp, C)[:, :hd].reshape(C, C, 1, 1, 1)
        grads[self.v_proj.weight] = dwkv[HP:].view(h, hdp, C)[:, :hd].reshape(C, C, 1, 1, 1)
        grads[self.k_proj.bias] = bkv[:HP].view(h, hdp)[:, :hd].reshape(C)
        grads[self.v_proj.bias] = bkv[HP:].view(h, hdp)[:, :hd].reshape(C)
        d_kv = dkv_dst if dkv_dst is not None else Blocked(n, C, Z, Y, X, False, dev)
        K.conv3d(dkvb, W["kv"], K.a_chunk_table(dkvb, [0], [2 * HP], False), d_kv.t, _lib.OUT_BLOCKED_BF16, dst_cbt=d_kv.cbt,
                 dst_cb_off=dkv_c0 // 8)
        return d_q, d_kv, grads

    def _bwd_weights(self):
        """dgrad-form operands of the three projections (column = input channel, K = output channel in the padded head
        layout), packed by the repack kernel from the same padded matrices the forward uses."""
        C, h, hd = self.in_channels, self.num_heads, self.head_dim
        ver = tuple(p._version for p in self.parameters())
        c = self.__dict__.get("_bwd_cache")
        if c is not None and c[0] == ver:
            return c[1]
        hdp = max(16, (hd + 15) // 16 * 16)
        r2 = lambda conv: conv.weight.detach().float().reshape(C, C)
        wq, _ = _pad_heads_out(r2(self.q_proj), self.q_proj.bias, h, hd, hdp)
        wk, _ = _pad_heads_out(r2(self.k_proj), self.k_proj.bias, h, hd, hdp)
        wv, _ = _pad_heads_out(r2(self.v_proj), self.v_proj.bias, h, hd, hdp)
        wkv = torch.cat([wk, wv])
        wo = torch.zeros((C, h * hdp), dtype=torch.float32, device=wq.device)
        wo.view(C, h, hdp)[:, :, :hd] = r2(self.out_proj).view(C, h, hd)
        f5 = lambda w: w.reshape(w.shape[0], w.shape[1], 1, 1, 1).contiguous()
        keep = {"q": f5(wq), "kv": f5(wkv), "o": f5(wo)}
        packs = {k_: K.PackPlan.k1_dgrad(v_).run() for k_, v_ in keep.items()}
        packs["_keep"] = keep
        self.__dict__["_bwd_cache"] = (ver, packs)
        return packs

    def forward(self, query_features: torch.Tensor, key_value_features: torch.Tensor) -> torch.Tensor:
        _require_cuda(query_features)
        B, C, Z, Y, X = query_features.shape
        if C % 16:
            raise NotImplementedError("CrossAttentionFusion kernels need channels % 16 == 0")
        if _wants_grad(self, query_features) or (torch.is_grad_enabled() and key_value_features.requires_grad):
            params = [self.q_proj.weight, self.q_proj.bias, self.k_proj.weight, self.k_proj.bias, self.v_proj.weight,
                      self.v_proj.bias, self.out_proj.weight, self.out_proj.bias]
            return _CrossAttentionFunction.apply(self, query_features, key_value_features, *params)
        with torch.no_grad():
            dev = query_features.device
            q_in = Blocked(B, C, Z, Y, X, False, dev)
            kv_in = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw(query_features.contiguous().float(), q_in)
            K.pack_ncdhw(key_value_features.contiguous().float(), kv_in)
            dst = Blocked(B, C, Z, Y, X, False, dev)
            self.forward_blocked(q_in, kv_in, dst)
            return dst.to_ncdhw()


class _CrossAttentionFunction(torch.autograd.Function):
    """CrossAttentionFusion.forward with the whole backward in the sm_100a kernels (bf16 operands, fp32 accumulation)."""

    @staticmethod
    def forward(ctx, module, q_feat, kv_feat, *params):
        dev = q_feat.device
        B, C, Z, Y, X = q_feat.shape
        with torch.no_grad():
            q_in = Blocked(B, C, Z, Y, X, False, dev)
            kv_in = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw(q_feat.detach().contiguous().float(), q_in)
            K.pack_ncdhw(kv_feat.detach().contiguous().float(), kv_in)
            dst = Blocked(B, C, Z, Y, X, False, dev)
            saved = {}
            module.forward_blocked(q_in, kv_in, dst, save=saved)
        ctx.module, ctx.saved, ctx.params = module, saved, params
        ctx.dtypes = (q_feat.dtype, kv_feat.dtype)
        return dst.to_ncdhw()

    @staticmethod
    def backward(ctx, g):
        m, saved = ctx.module, ctx.saved
        q_in = saved["q_in"]
        with torch.no_grad():
            g_out = Blocked(q_in.n_img, m.in_channels, q_in.Z, q_in.Y, q_in.X, False, g.device)
            K.pack_ncdhw(g.contiguous().float(), g_out)
            d_q, d_kv, grads = m.backward_blocked(saved, g_out)
            gq = d_q.to_ncdhw().to(ctx.dtypes[0]) if ctx.needs_input_grad[1] else None
            gkv = d_kv.to_ncdhw().to(ctx.dtypes[1]) if ctx.needs_input_grad[2] else None
            gp = [grads[p].to(p.dtype).view_as(p) if (p.requires_grad and p in grads) else None for p in ctx.params]
        return (None, gq, gkv, *gp)


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
        B, C, Z, Y, X = features_1.shape
        if _wants_grad(self, features_1) or (torch.is_grad_enabled() and features_2.requires_grad):
            params = [p for p in self.parameters()]
            return _BidirectionalFunction.apply(self, features_1, features_2, *params)
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


class _BidirectionalFunction(torch.autograd.Function):
    """BidirectionalCrossAttention.forward (reference attention_fusion.py:193-216) with the backward in the kernels: the two
    CrossAttentionFusion backwards, the 1x1 fusion conv's dgrad / wgrad and the InstanceNorm + ReLU backward."""

    @staticmethod
    def forward(ctx, module, f1_t, f2_t, *params):
        dev = f1_t.device
        B, C, Z, Y, X = f1_t.shape
        with torch.no_grad():
            f1 = Blocked(B, C, Z, Y, X, False, dev)
            f2 = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw(f1_t.detach().contiguous().float(), f1)
            K.pack_ncdhw(f2_t.detach().contiguous().float(), f2)
            cat = Blocked(B, 2 * C, Z, Y, X, False, dev)
            s1, s2 = {}, {}
            module.cross_attn_1to2.forward_blocked(f1, f2, cat, 0, save=s1)
            module.cross_attn_2to1.forward_blocked(f2, f1, cat, C, save=s2)
            conv = module.fusion[0]
            pw = K.PackPlan.forward(conv.weight, None, False, [C, C], use_bias=False).run()   # bias cancelled by the norm
            a_cb = K.a_chunk_table(cat, [0, C], [C, C], False)
            tile = K.plan_conv_norm((X, Y, Z), B, pw, False, a_cb)
            raw = torch.empty((B, C // 8, Z, Y, X, 8), dtype=torch.bfloat16, device=dev)
            stats = torch.empty(B * tile.tiles_per_img * C * 2, dtype=torch.float32, device=dev)
            mr = torch.empty((B, C, 2), dtype=torch.float32, device=dev)
            K.conv3d(cat, pw, a_cb, raw, _lib.OUT_BLOCKED_BF16, stats=stats, dst_cbt=C // 8, tile=tile)
            K.instnorm_finalize(stats, B, tile.tiles_per_img, C, Z * Y * X, mr, eps=module.fusion[1].eps)
            out = Blocked(B, C, Z, Y, X, False, dev)
            K.instnorm_act_apply(raw, False, mr, B, C, Z, Y, X, out, 0, 0.0)
        ctx.module, ctx.saved, ctx.params = module, (s1, s2, cat, raw, mr), params
        ctx.dtypes = (f1_t.dtype, f2_t.dtype)
        return out.to_ncdhw()

    @staticmethod
    def backward(ctx, g):
        m = ctx.module
        s1, s2, cat, raw, mr = ctx.saved
        B, C, Z, Y, X = cat.n_img, m.out_channels, cat.Z, cat.Y, cat.X
        dev = g.device
        conv = m.fusion[0]
        with torch.no_grad():
            g_out = Blocked(B, C, Z, Y, X, False, dev)
            K.pack_ncdhw(g.contiguous().float(), g_out)
            draw = Blocked(B, C, Z, Y, X, False, dev)
            K.instnorm_act_bwd(raw, mr, B, C, Z, Y, X, g_out, 0, 1.0, None, 0, draw.t, slope=0.0)
            grads = {conv.weight: K.conv3d_wgrad(cat, [(0, C), (C, C)], draw.t, draw.cbt, 0, C, 1, conv.weight.shape)}
            if conv.bias is not None:
                grads[conv.bias] = torch.zeros_like(conv.bias, dtype=torch.float32)    # cancelled by the InstanceNorm
            d_cat = Blocked(B, 2 * C, Z, Y, X, False, dev)
            pwd = K.PackPlan.k1_dgrad(conv.weight.detach()).run()
            K.conv3d(draw, pwd, K.a_chunk_table(draw, [0], [C], False), d_cat.t, _lib.OUT_BLOCKED_BF16, dst_cbt=d_cat.cbt)
            # d f1 = dq(1->2) + dkv(2->1);  d f2 = dkv(1->2) + dq(2->1): each pair lands side by side, then one add
            sum1 = Blocked(B, 2 * C, Z, Y, X, False, dev)
            sum2 = Blocked(B, 2 * C, Z, Y, X, False, dev)
            _, _, g1 = m.cross_attn_1to2.backward_blocked(s1, d_cat, 0, dq_dst=sum1, dq_c0=0, dkv_dst=sum2, dkv_c0=0)
            _, _, g2 = m.cross_attn_2to1.backward_blocked(s2, d_cat, C, dq_dst=sum2, dq_c0=C, dkv_dst=sum1, dkv_c0=C)
            grads.update(g1)
            grads.update(g2)
            d1 = Blocked(B, C, Z, Y, X, False, dev)
            d2 = Blocked(B, C, Z, Y, X, False, dev)
            K.modality_combine(sum1, 2, C, d1, 0, None, 1.0)
            K.modality_combine(sum2, 2, C, d2, 0, None, 1.0)
            gf1 = d1.to_ncdhw().to(ctx.dtypes[0]) if ctx.needs_input_grad[1] else None
            gf2 = d2.to_ncdhw().to(ctx.dtypes[1]) if ctx.needs_input_grad[2] else None
            gp = [grads[p].to(p.dtype).view_as(p) if (p.requires_grad and p in grads) else None for p in ctx.params]
        return (None, gf1, gf2, *gp)


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
