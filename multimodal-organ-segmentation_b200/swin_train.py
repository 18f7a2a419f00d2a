"""Training path of SwinUNETR (BASELINE.json configs[3]: forward/backward bf16): every op of MONAI's SwinUNETR.forward as a
torch.autograd.Function whose forward AND backward are C-ABI kernel launches (bf16 operands, fp32 accumulation, fp32 token
residual stream and weight gradients).  Autograd only provides the tape and sums the gradients of tensors with several
consumers (residual stream, skip connections); the reference reaches the same ops through loss.backward()
(src/trainer/trainer.py:243) over monai.networks.nets.SwinUNETR (src/models/backbones/swin_unetr.py:80-117).

Tensors are plain torch tensors in the blocked layout [n, C/8, Z, Y, X, 8]: "16" = bf16 (GEMM operands, activations),
"32" = fp32 (token residual stream and its gradient).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import ctypes as C

import torch
from torch.autograd import Function

from . import _lib
from . import kernels as K
from .kernels import Blocked, _call, _ptr, _stream
from .train_engine import _wrap

Tensor = torch.Tensor
LRELU_SLOPE = 0.01
_PLANS: Dict[Tuple, K.PackPlan] = {}


_FRESH: Dict[Tuple, int] = {}      # plan key -> parameter version its packed operand was derived from
_BATCHES: Dict[int, Tuple] = {}


def _plan(param: Tensor, variant, make) -> K.PackedConv:
    """Packed operand of `param`: re-derived by the step's batched repack (`_repack_all`), or by its own launch when the
    plan is new."""
    key = (id(param), variant)
    plan = _PLANS.get(key)
    if plan is None or plan.weight.data_ptr() != param.data_ptr():
        if len(_PLANS) > 4096:
            _PLANS.clear()
            _BATCHES.clear()
        plan = _PLANS[key] = make()
        _BATCHES.clear()
        _FRESH.pop(key, None)    # (an id can be re-used by a new parameter: the new plan has not run yet)
    if _FRESH.get(key) == plan.weight._version:      # (detached views share the parameter's version counter)
        return plan.pc
    _FRESH[key] = plan.weight._version
    return plan.run()


def _repack_all(net) -> None:
    """One launch re-derives every packed operand of `net` from its live parameters (kernels.PackBatch)."""
    ids = {id(p) for p in net.parameters()}
    ent = _BATCHES.get(id(net))
    if ent is None or not ent[0].valid():
        live = {k: p for k, p in _PLANS.items() if k[0] in ids}
        if not live:
            return
        ent = _BATCHES[id(net)] = (K.PackBatch(list(live.values())), list(live.keys()))
    if all(_FRESH.get(k) == p.weight._version for k, p in zip(ent[1], ent[0].plans)):
        return                                       # parameters unchanged since the last repack (gradient accumulation)
    ent[0].run()
    for k, p in zip(ent[1], ent[0].plans):
        _FRESH[k] = p.weight._version


def _B(t: Tensor) -> Blocked:
    n, cb, Z, Y, X, _ = t.shape
    assert t.dtype == torch.bfloat16 and t.is_contiguous()
    return _wrap(t, n, cb * 8, Z, Y, X)


def _new16(n, channels, dims, device) -> Tensor:
    return torch.empty((n, channels // 8, dims[0], dims[1], dims[2], 8), dtype=torch.bfloat16, device=device)


def _new32(n, channels, dims, device) -> Tensor:
    return torch.empty((n, channels // 8, dims[0], dims[1], dims[2], 8), dtype=torch.float32, device=device)


def _w5(weight: Tensor) -> Tensor:
    return weight.detach().view(weight.shape[0], weight.shape[1], 1, 1, 1) if weight.dim() == 2 else weight.detach()


def _channel_sums(t16: Tensor) -> Tensor:
    b = _B(t16)
    return (K.channel_mean(b, 0, b.channels) * float(b.nvox)).sum(0)


def _cast16(t32: Tensor) -> Tensor:
    """fp32 blocked -> bf16 blocked (the pass-through mode of the LayerNorm backward kernel)."""
    n, cb, Z, Y, X, _ = t32.shape
    out = torch.empty(t32.shape, dtype=torch.bfloat16, device=t32.device)
    _call("mmseg_swin_ln_bwd", None, None, None, None, _ptr(t32), None, _ptr(out), n, cb, Z * Y * X, _stream())
    return out


# ------------------------------------------------------------------------------------------------ linear / conv ops
class LinearFn(Function):
    """nn.Linear over tokens = 1x1x1 GEMM on the tcgen05 conv kernel; out32: fp32 output (the new residual stream after a
    patch-merging reduction)."""

    @staticmethod
    def forward(ctx, x16, weight, bias, out32: bool):
        n, cb, Z, Y, X, _ = x16.shape
        cout, cin = weight.shape
        w5 = _w5(weight)
        pw = _plan(weight, "lin", lambda: K.PackPlan.forward(w5, bias.detach() if bias is not None else None, False))
        xb = _B(x16)
        a_cb = K.a_chunk_table(xb, [0], [cin], False)
        y = (_new32 if out32 else _new16)(n, cout, (Z, Y, X), x16.device)
        K.conv3d(xb, pw, a_cb, y, _lib.OUT_BLOCKED_F32 if out32 else _lib.OUT_BLOCKED_BF16, dst_cbt=cout // 8)
        ctx.save_for_backward(x16, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x16, weight = ctx.saved_tensors
        cout, cin = weight.shape
        dy16 = dy.contiguous() if dy.dtype == torch.bfloat16 else _cast16(dy.contiguous())
        n, _, Z, Y, X, _ = x16.shape
        w5 = _w5(weight)
        dx = None
        if ctx.needs_input_grad[0]:
            pwd = _plan(weight, "lin_dgrad", lambda: K.PackPlan.k1_dgrad(w5))
            dyb = _B(dy16)
            dx = _new16(n, cin, (Z, Y, X), x16.device)
            K.conv3d(dyb, pwd, K.a_chunk_table(dyb, [0], [cout], False), dx, _lib.OUT_BLOCKED_BF16, dst_cbt=cin // 8)
        dw = K.conv3d_wgrad(_B(x16), [(0, cin)], dy16, cout // 8, 0, cout, 1, (cout, cin, 1, 1, 1)).view(cout, cin)
        db = _channel_sums(dy16) if ctx.has_bias else None
        return dx, dw, db, None


class ConvStatsFn(Function):
    """Conv3d (k3 pad 1 / k1, no bias) -> raw bf16 output + the InstanceNorm (mean, rstd) table from the conv epilogue."""

    @staticmethod
    def forward(ctx, x16, weight, segs):
        n, cb, Z, Y, X, _ = x16.shape
        cout, ks = weight.shape[0], weight.shape[2]
        seg_ch = tuple(s[1] for s in segs)
        pw = _plan(weight, ("conv", seg_ch), lambda: K.PackPlan.forward(weight.detach(), None, False, seg_ch, use_bias=False))
        xb = _B(x16)
        a_cb = K.a_chunk_table(xb, [s[0] for s in segs], list(seg_ch), False)
        tile = K.plan_conv_norm((X, Y, Z), n, pw, False, a_cb)
        raw = _new16(n, cout, (Z, Y, X), x16.device)
        stats = torch.empty(n * tile.tiles_per_img * cout * 2, dtype=torch.float32, device=x16.device)
        K.conv3d(xb, pw, a_cb, raw, _lib.OUT_BLOCKED_BF16, stats=stats, dst_cbt=cout // 8, tile=tile)
        mr = torch.empty((n, cout, 2), dtype=torch.float32, device=x16.device)
        K.instnorm_finalize(stats, n, tile.tiles_per_img, cout, Z * Y * X, mr)
        ctx.save_for_backward(x16, weight)
        ctx.segs = list(segs)
        ctx.mark_non_differentiable(mr)
        return raw, mr

    @staticmethod
    def backward(ctx, draw, _dmr):
        x16, weight = ctx.saved_tensors
        segs = ctx.segs
        cout, ks = weight.shape[0], weight.shape[2]
        n, cb, Z, Y, X, _ = x16.shape
        draw = draw.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            mk = (lambda: K.PackPlan.dgrad(weight.detach())) if ks == 3 else (lambda: K.PackPlan.k1_dgrad(weight.detach()))
            pwd = _plan(weight, "conv_dgrad", mk)
            db_ = _B(draw)
            dx = torch.zeros_like(x16) if segs[0][0] != 0 or pwd.n_out != cb * 8 else torch.empty_like(x16)
            K.conv3d(db_, pwd, K.a_chunk_table(db_, [0], [cout], False), dx, _lib.OUT_BLOCKED_BF16, dst_cbt=cb,
                     dst_cb_off=segs[0][0] // 8)
        dw = K.conv3d_wgrad(_B(x16), segs, draw, cout // 8, 0, cout, ks, tuple(weight.shape))
        return dx, dw, None


class NormActFn(Function):
    """InstanceNorm3d(affine=False) apply + LeakyReLU(slope)."""

    @staticmethod
    def forward(ctx, raw, mr, slope: float):
        n, cb, Z, Y, X, _ = raw.shape
        y = torch.empty_like(raw)
        K.instnorm_act_apply(raw, False, mr, n, cb * 8, Z, Y, X, _B(y), 0, slope)
        ctx.save_for_backward(raw, mr)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        raw, mr = ctx.saved_tensors
        n, cb, Z, Y, X, _ = raw.shape
        draw = torch.empty_like(raw)
        K.instnorm_act_bwd(raw, mr, n, cb * 8, Z, Y, X, _B(dy.contiguous()), 0, 1.0, None, 0, draw, ctx.slope)
        return draw, None, None


class ResTailFn(Function):
    """UnetResBlock tail: y = LeakyReLU(IN(raw2) + (IN(raw3) | x))."""

    @staticmethod
    def forward(ctx, raw2, mr2, res, mr3):
        n, cb, Z, Y, X, _ = raw2.shape
        y = torch.empty_like(raw2)
        K.instnorm_residual_act(raw2, False, mr2, res, False, mr3, cb, 0, _B(y), 0, n, cb * 8, Z * Y * X, LRELU_SLOPE)
        ctx.save_for_backward(raw2, mr2, res, mr3 if mr3 is not None else mr2, y)
        ctx.has_conv3 = mr3 is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        raw2, mr2, res, mr3, y = ctx.saved_tensors
        n, cb, Z, Y, X, _ = raw2.shape
        g = torch.empty_like(raw2)
        _call("mmseg_lrelu_mask_mul", _ptr(y), _ptr(dy.contiguous()), _ptr(g), g.numel(), LRELU_SLOPE, _stream())
        gb = _B(g)
        draw2 = torch.empty_like(raw2)
        K.instnorm_act_bwd(raw2, mr2, n, cb * 8, Z, Y, X, gb, 0, 1.0, None, 0, draw2, 1.0)   # slope 1: identity activation
        if ctx.has_conv3:
            dres = torch.empty_like(res)
            K.instnorm_act_bwd(res, mr3, n, cb * 8, Z, Y, X, gb, 0, 1.0, None, 0, dres, 1.0)
        else:
            dres = g
        return draw2, None, dres, None


class ConvTransposeFn(Function):
    """ConvTranspose3d(k2, s2, no bias) = 1x1x1 GEMM + pixel shuffle."""

    @staticmethod
    def forward(ctx, x16, weight):
        n, cb, Z, Y, X, _ = x16.shape
        cin, f = weight.shape[0], weight.shape[1]
        pw = _plan(weight, "convt", lambda: K.PackPlan.forward(weight.detach(), None, False, None, transposed=True))
        xb = _B(x16)
        y = _new16(n, f, (2 * Z, 2 * Y, 2 * X), x16.device)
        K.conv3d(xb, pw, K.a_chunk_table(xb, [0], [cin], False), y, _lib.OUT_CONVT_K2S2, dst_cbt=f // 8)
        ctx.save_for_backward(x16, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x16, weight = ctx.saved_tensors
        n, cb, Z, Y, X, _ = x16.shape
        cin, f = weight.shape[0], weight.shape[1]
        dyu = _new16(n, 8 * f, (Z, Y, X), x16.device)
        K.unshuffle_k2s2(_B(dy.contiguous()), 0, f, dyu)
        dw = K.conv3d_wgrad(_B(x16), [(0, cin)], dyu, f, 0, 8 * f, 1, tuple(weight.shape), transposed=True)
        pwd = _plan(weight, "convt_dgrad", lambda: K.PackPlan.convt_dgrad(weight.detach()))
        db_ = _B(dyu)
        dx = torch.empty_like(x16)
        K.conv3d(db_, pwd, K.a_chunk_table(db_, [0], [8 * f], False), dx, _lib.OUT_BLOCKED_BF16, dst_cbt=cb)
        return dx, dw


class OutConvFn(Function):
    """UnetOutBlock: Conv3d(F, classes, 1) with bias -> NCDHW fp32 logits."""

    @staticmethod
    def forward(ctx, x16, weight, bias):
        n, cb, Z, Y, X, _ = x16.shape
        logits = torch.empty((n, weight.shape[0], Z, Y, X), dtype=torch.float32, device=x16.device)
        K.conv1x1_logits(_B(x16), 0, cb * 8, weight.detach(), bias.detach() if bias is not None else None, logits)
        ctx.save_for_backward(x16, weight)
        ctx.has_bias = bias is not None
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x16, weight = ctx.saved_tensors
        n, cb, Z, Y, X, _ = x16.shape
        k, cin = weight.shape[0], weight.shape[1]
        kp = (k + 15) // 16 * 16
        dl = Blocked(n, kp, Z, Y, X, False, x16.device)
        K.pack_ncdhw(dlogits.contiguous().float(), dl)
        dw = K.conv3d_wgrad(_B(x16), [(0, cin)], dl.t, dl.cbt, 0, k, 1, tuple(weight.shape))
        db = (K.channel_mean(dl, 0, kp) * float(dl.nvox)).sum(0)[:k] if ctx.has_bias else None
        pwd = _plan(weight, "k1_dgrad", lambda: K.PackPlan.k1_dgrad(weight.detach()))
        dx = torch.empty_like(x16)
        K.conv3d(dl, pwd, K.a_chunk_table(dl, [0], [k], False), dx, _lib.OUT_BLOCKED_BF16, dst_cbt=cb)
        return dx, dw, db


# ------------------------------------------------------------------------------------------------ token ops
class PatchEmbedFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        n, cin, Z2, Y2, X2 = x.shape
        F_ = weight.shape[0]
        xs = _new32(n, F_, (Z2 // 2, Y2 // 2, X2 // 2), x.device)
        K.swin_patch_embed(x, weight.detach(), bias.detach() if bias is not None else None, xs)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return xs

    @staticmethod
    def backward(ctx, dxs):
        x, weight = ctx.saved_tensors
        n, cin, Z2, Y2, X2 = x.shape
        F_ = weight.shape[0]
        K1 = cin * 8 + 1
        n_chunks = 296
        partial = torch.empty((n_chunks, F_, K1), dtype=torch.float32, device=x.device)
        _call("mmseg_swin_patch_embed_wgrad", _ptr(x), _ptr(dxs.contiguous()), _ptr(partial), n_chunks, n, cin, F_, Z2 // 2,
              Y2 // 2, X2 // 2, _stream())
        g = partial.sum(0)
        dw = g[:, :K1 - 1].reshape(F_, cin, 2, 2, 2)
        return None, dw, (g[:, K1 - 1].contiguous() if ctx.has_bias else None)


def _ln_fwd(xs_in, add16, gamma, beta, want_ln: bool, eps: float):
    n, cb, Z, Y, X, _ = xs_in.shape
    vox = Z * Y * X
    xs_out = torch.empty_like(xs_in) if add16 is not None else xs_in
    ln = torch.empty(xs_in.shape, dtype=torch.bfloat16, device=xs_in.device) if want_ln else None
    stats = torch.empty((n * vox, 2), dtype=torch.float32, device=xs_in.device) if want_ln else None
    _call("mmseg_swin_ln_fwd_train", _ptr(xs_in), _ptr(add16), _ptr(gamma), _ptr(beta), _ptr(xs_out), _ptr(ln), _ptr(stats), n, cb,
          vox, eps, _stream())
    return xs_out, ln, stats


def _ln_bwd(xs, stats, dln16, gamma, dxs_in, want32: bool, want16: bool):
    n, cb, Z, Y, X, _ = (xs if xs is not None else dxs_in).shape
    ref = xs if xs is not None else dxs_in
    d32 = torch.empty(ref.shape, dtype=torch.float32, device=ref.device) if want32 else None
    d16 = torch.empty(ref.shape, dtype=torch.bfloat16, device=ref.device) if want16 else None
    _call("mmseg_swin_ln_bwd", _ptr(xs), _ptr(stats), _ptr(dln16), _ptr(gamma), _ptr(dxs_in), _ptr(d32), _ptr(d16), n, cb,
          Z * Y * X, _stream())
    return d32, d16


def _ln_param_grad(xs, stats, dln16):
    n, cb, Z, Y, X, _ = xs.shape
    n_chunks = max(1, min(64, (n * Z * Y * X + 2047) // 2048))
    partial = torch.empty((n_chunks, cb * 8, 2), dtype=torch.float32, device=xs.device)
    _call("mmseg_swin_ln_param_grad", _ptr(xs), _ptr(stats), _ptr(dln16), _ptr(partial), n_chunks, n, cb, Z * Y * X, _stream())
    g = partial.sum(0)
    return g[:, 0].contiguous(), g[:, 1].contiguous()


class LayerNormFn(Function):
    """ln16 = LayerNorm_C(xs) (optional affine): norm1 of a stage's first block, proj_out of the hidden states, the
    LayerNorm(8C) of PatchMerging."""

    @staticmethod
    def forward(ctx, xs, gamma, beta, eps: float):
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        _, ln, stats = _ln_fwd(xs, None, g, b, True, eps)
        ctx.save_for_backward(xs, stats, g if g is not None else stats)
        ctx.affine = gamma is not None
        return ln

    @staticmethod
    def backward(ctx, dln):
        xs, stats, g = ctx.saved_tensors
        gamma = g if ctx.affine else None
        dln = dln.contiguous()
        dxs, _ = _ln_bwd(xs, stats, dln, gamma, None, True, False)
        dg = db = None
        if ctx.affine:
            dg, db = _ln_param_grad(xs, stats, dln)
        return dxs, dg, db, None


class AddLayerNormFn(Function):
    """xs_out = xs_in + y16;  ln16 = LayerNorm_C(xs_out) * gamma + beta — the residual add fused with the next norm."""

    @staticmethod
    def forward(ctx, xs_in, y16, gamma, beta, eps: float):
        g, b = gamma.detach(), beta.detach()
        xs_out, ln, stats = _ln_fwd(xs_in, y16, g, b, True, eps)
        ctx.save_for_backward(xs_out, stats, g)
        return xs_out, ln

    @staticmethod
    def backward(ctx, dxs_out, dln):
        xs_out, stats, gamma = ctx.saved_tensors
        dln = dln.contiguous()
        d32, d16 = _ln_bwd(xs_out, stats, dln, gamma, dxs_out.contiguous() if dxs_out is not None else None, True, True)
        dg, db = _ln_param_grad(xs_out, stats, dln)
        return d32, d16, dg, db, None


class AddFn(Function):
    """xs_out = xs_in + y16 (the last residual add of a stage)."""

    @staticmethod
    def forward(ctx, xs_in, y16):
        xs_out, _, _ = _ln_fwd(xs_in, y16, None, None, False, 1e-5)
        return xs_out

    @staticmethod
    def backward(ctx, dxs_out):
        dxs_out = dxs_out.contiguous()
        _, d16 = _ln_bwd(None, None, None, None, dxs_out, False, True)
        return dxs_out, d16


class GeluFn(Function):
    @staticmethod
    def forward(ctx, h16, ident):
        n, cb, Z, Y, X, _ = h16.shape
        y = torch.empty_like(h16)
        K.instnorm_act_apply(h16, False, ident, n, cb * 8, Z, Y, X, _B(y), 0, gelu=True)
        ctx.save_for_backward(h16)
        return y

    @staticmethod
    def backward(ctx, dy):
        (h16,) = ctx.saved_tensors
        dx = torch.empty_like(h16)
        _call("mmseg_gelu_bwd", _ptr(h16), _ptr(dy.contiguous()), _ptr(dx), h16.numel(), _stream())
        return dx, None


def _attn_args(qkv16, out16, table, qkv_bias, heads, window, shift, lse):
    n, cb3, Z, Y, X, _ = qkv16.shape
    a = _lib.SwinAttnArgs()
    a.qkv, a.out, a.table = qkv16.data_ptr(), out16.data_ptr(), table.data_ptr()
    a.qkv_bias = qkv_bias.data_ptr() if qkv_bias is not None else None
    a.n_img, a.D, a.H, a.W = n, Z, Y, X
    for i in range(3):
        a.window[i], a.shift[i] = int(window[i]), int(shift[i])
    a.heads, a.head_dim = heads, 16
    a.qkv_cbt, a.out_cbt, a.out_cb_off = cb3, cb3 // 3, 0
    a.scale, a.elem_fmt = 0.25, _lib.FMT_BF16
    a.lse = lse.data_ptr() if lse is not None else None
    return a


def _n_windows(dims, window) -> int:
    n = 1
    for d, w in zip(dims, window):
        ws = d if d <= w else w
        n *= -(-d // ws)
    return n


class WindowAttentionFn(Function):
    @staticmethod
    def forward(ctx, qkv16, table, qkv_bias, heads: int, window, shift):
        n, cb3, Z, Y, X, _ = qkv16.shape
        out = torch.empty((n, cb3 // 3, Z, Y, X, 8), dtype=torch.bfloat16, device=qkv16.device)
        nw = _n_windows((Z, Y, X), window)
        lse = torch.empty((n, heads, nw, 352), dtype=torch.float32, device=qkv16.device)
        tb = table.detach()
        qb = qkv_bias.detach() if qkv_bias is not None else None
        a = _attn_args(qkv16, out, tb, qb, heads, window, shift, lse)
        _call("mmseg_swin_window_attention", C.byref(a), _stream())
        ctx.save_for_backward(qkv16, out, lse, tb, qb if qb is not None else tb)
        ctx.cfg = (heads, tuple(window), tuple(shift), qkv_bias is not None, nw)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv16, out, lse, table, qb = ctx.saved_tensors
        heads, window, shift, has_bias, nw = ctx.cfg
        n = qkv16.shape[0]
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv16)
        dtab = torch.empty((n * nw, heads, table.shape[0]), dtype=torch.float32, device=qkv16.device)
        dbias = torch.empty((n * nw, heads, 32), dtype=torch.float32, device=qkv16.device)
        a = _attn_args(qkv16, out, table, qb if has_bias else None, heads, window, shift, None)
        _call("mmseg_swin_window_attention_bwd", C.byref(a), _ptr(dout), dout.shape[1], 0, _ptr(lse), _ptr(dqkv), _ptr(dtab),
              _ptr(dbias), _stream())
        dtable = dtab.sum(0).t().contiguous()                     # [table rows, heads]
        dqb = None
        if has_bias:
            pad = dbias.sum(0)                                    # [heads, 32]: dk | dv of the zero-padded tokens
            Cc = heads * 16
            dqb = torch.zeros(3 * Cc, dtype=torch.float32, device=qkv16.device)
            dqb[Cc:2 * Cc] = pad[:, :16].reshape(-1)
            dqb[2 * Cc:] = pad[:, 16:].reshape(-1)
        return dqkv, dtable, dqb, None, None, None


class MergeGatherFn(Function):
    @staticmethod
    def forward(ctx, xs):
        n, cb, Z, Y, X, _ = xs.shape
        cat = torch.empty((n, 8 * cb, Z // 2, Y // 2, X // 2, 8), dtype=torch.float32, device=xs.device)
        _call("mmseg_swin_merge_gather", _ptr(xs), _ptr(cat), n, cb, Z, Y, X, _stream())
        ctx.shape = tuple(xs.shape)
        return cat

    @staticmethod
    def backward(ctx, dcat):
        n, cb, Z, Y, X, _ = ctx.shape
        dxs = torch.empty(ctx.shape, dtype=torch.float32, device=dcat.device)
        _call("mmseg_swin_merge_scatter", _ptr(dcat.contiguous()), _ptr(dxs), n, cb, Z, Y, X, _stream())
        return dxs


# ------------------------------------------------------------------------------------------------ the model
def _res_block(blk, x16: Tensor, segs) -> Tensor:
    """MONAI UnetResBlock on blocked bf16 tensors."""
    raw1, mr1 = ConvStatsFn.apply(x16, blk.conv1.conv.weight, segs)
    h = NormActFn.apply(raw1, mr1, LRELU_SLOPE)
    cout = blk.conv2.conv.weight.shape[0]
    raw2, mr2 = ConvStatsFn.apply(h, blk.conv2.conv.weight, [(0, cout)])
    conv3 = getattr(blk, "conv3", None)
    if conv3 is not None:
        raw3, mr3 = ConvStatsFn.apply(x16, conv3.conv.weight, segs)
        return ResTailFn.apply(raw2, mr2, raw3, mr3)
    return ResTailFn.apply(raw2, mr2, x16, None)


def swin_unetr_train_forward(net, x: Tensor) -> Tensor:
    """SwinUNETR.forward with a tape: logits NCDHW fp32 that autograd can differentiate w.r.t. every parameter."""
    _lib.require_device()
    x = x.contiguous().float()
    n, cin, Z, Y, X = x.shape
    if any(d % 32 for d in (Z, Y, X)):
        raise NotImplementedError(f"SwinUNETR needs spatial sizes divisible by 32, got {(Z, Y, X)}")
    F_ = net.feature_size
    dev = x.device
    win = tuple(net.window_size)
    vit = net.swinViT
    pe = vit.patch_embed.proj
    _repack_all(net)
    xs = PatchEmbedFn.apply(x, pe.weight, pe.bias)
    hidden = [LayerNormFn.apply(xs, None, None, 1e-5)]
    for s in range(4):
        layer = getattr(vit, f"layers{s + 1}")[0]
        Cc = F_ << s
        heads = layer.blocks[0].attn.num_heads
        ident = torch.zeros((n, 4 * Cc, 2), dtype=torch.float32, device=dev)
        ident[:, :, 1] = 1.0
        blk0 = layer.blocks[0]
        ln = LayerNormFn.apply(xs, blk0.norm1.weight, blk0.norm1.bias, blk0.norm1.eps)
        depth = len(layer.blocks)
        for j, blk in enumerate(layer.blocks):
            shift = (0, 0, 0) if j % 2 == 0 else tuple(w // 2 for w in win)
            qkv = LinearFn.apply(ln, blk.attn.qkv.weight, blk.attn.qkv.bias, False)
            att = WindowAttentionFn.apply(qkv, blk.attn.relative_position_bias_table, blk.attn.qkv.bias, heads, win, shift)
            y = LinearFn.apply(att, blk.attn.proj.weight, blk.attn.proj.bias, False)
            xs, ln = AddLayerNormFn.apply(xs, y, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            h = LinearFn.apply(ln, blk.mlp.linear1.weight, blk.mlp.linear1.bias, False)
            h = GeluFn.apply(h, ident)
            y = LinearFn.apply(h, blk.mlp.linear2.weight, blk.mlp.linear2.bias, False)
            if j + 1 < depth:
                nxt = layer.blocks[j + 1]
                xs, ln = AddLayerNormFn.apply(xs, y, nxt.norm1.weight, nxt.norm1.bias, nxt.norm1.eps)
            else:
                xs = AddFn.apply(xs, y)
        ds = layer.downsample
        cat = MergeGatherFn.apply(xs)
        mg = LayerNormFn.apply(cat, ds.norm.weight, ds.norm.bias, ds.norm.eps)
        xs = LinearFn.apply(mg, ds.reduction.weight, None, True)
        hidden.append(LayerNormFn.apply(xs, None, None, 1e-5))

    # UNETR encoder / decoder
    xin = Blocked(n, (cin + 15) // 16 * 16, Z, Y, X, False, dev)
    xin.t.zero_()
    K.pack_ncdhw(x, xin)
    enc0 = _res_block(net.encoder1.layer, xin.t, [(0, cin)])
    enc1 = _res_block(net.encoder2.layer, hidden[0], [(0, F_)])
    enc2 = _res_block(net.encoder3.layer, hidden[1], [(0, 2 * F_)])
    enc3 = _res_block(net.encoder4.layer, hidden[2], [(0, 4 * F_)])
    cur = _res_block(net.encoder10.layer, hidden[4], [(0, 16 * F_)])
    for dec, skip, c in ((net.decoder5, hidden[3], 8 * F_), (net.decoder4, enc3, 4 * F_), (net.decoder3, enc2, 2 * F_),
                         (net.decoder2, enc1, F_), (net.decoder1, enc0, F_)):
        up = ConvTransposeFn.apply(cur, dec.transp_conv.conv.weight)
        cat = torch.cat([up, skip], dim=1)
        cur = _res_block(dec.conv_block, cat, [(0, c), (c, c)])
    oc = net.out.conv.conv
    return OutConvFn.apply(cur, oc.weight, oc.bias)
