"""GPU parity cases of the SwinUNETR path (swin.cu + swin_engine.py) vs the CPU oracle (oracle/swin_unetr.py, PARITY
UNPINNED: restated from MONAI's published algorithm, see its header) / torch fp64 on the same seeded inputs.
Plain assert-based functions, called by tests/test_gpu_swin.py (marker `gpu`) and tools/gpu_check.py."""
import torch
import torch.nn.functional as F

import mmseg_b200  # noqa: F401
from mmseg_b200 import _lib
from mmseg_b200 import kernels as K
from mmseg_b200.kernels import Blocked
from mmseg_b200.swin_engine import SWIN_MODES
from oracle import swin_unetr as O

DEV = "cuda"


def _blocked_f32(x):
    """NCDHW fp32 -> blocked fp32 [n, C/8, Z, Y, X, 8]."""
    n, c, Z, Y, X = x.shape
    return x.view(n, c // 8, 8, Z, Y, X).permute(0, 1, 3, 4, 5, 2).contiguous()


def _unblocked_f32(xb):
    n, cb, Z, Y, X, _ = xb.shape
    return xb.permute(0, 1, 5, 2, 3, 4).reshape(n, cb * 8, Z, Y, X)


def _rnd16(x, mode):
    return x.to(SWIN_MODES[mode].dtype).float()


def patch_embed_case(cin=2, feat=48, dims=(6, 10, 12), n=2):
    torch.manual_seed(1)
    x = torch.randn(n, cin, *[2 * d for d in dims], device=DEV)
    w = torch.randn(feat, cin, 2, 2, 2, device=DEV) * 0.3
    b = torch.randn(feat, device=DEV)
    xs = torch.empty((n, feat // 8, *dims, 8), device=DEV)
    K.swin_patch_embed(x, w, b, xs)
    ref = F.conv3d(x.double(), w.double(), b.double(), stride=2).float()
    err = (_unblocked_f32(xs) - ref).abs().max().item()
    print(f"[patch_embed] max|err| {err:.2e}")
    assert err < 1e-4


def layernorm_case(channels=96, dims=(5, 6, 7), n=2, mode="fp16", with_add=True, affine=True, c0=0, extra=0):
    torch.manual_seed(2)
    x = torch.randn(n, channels, *dims, device=DEV) * 2 + 0.5
    y = torch.randn(n, channels, *dims, device=DEV)
    g = torch.randn(channels, device=DEV) if affine else None
    be = torch.randn(channels, device=DEV) if affine else None
    xs = _blocked_f32(x)
    dst = Blocked(n, channels + c0 + extra, *dims, SWIN_MODES[mode], DEV)
    dst.t.zero_()
    vox = dims[0] * dims[1] * dims[2]
    K.swin_layernorm(xs, n, channels, vox, dst, c0, add=_blocked_f32(y) if with_add else None, gamma=g, beta=be)
    s = (x + y) if with_add else x
    ref = F.layer_norm(s.double().permute(0, 2, 3, 4, 1), [channels], None if g is None else g.double(),
                       None if be is None else be.double(), 1e-5).permute(0, 4, 1, 2, 3).float()
    got = dst.to_ncdhw(c0, channels)
    e_ln = (got - ref).abs().max().item()
    e_x = (_unblocked_f32(xs) - s).abs().max().item()
    tol = 4e-3 if mode == "fp16" else 3e-2
    print(f"[layernorm c{channels} add{int(with_add)} affine{int(affine)}] LN err {e_ln:.2e} (tol {tol}), residual err {e_x:.2e}")
    assert e_ln < tol * max(1.0, ref.abs().max().item() / 4) and e_x < 1e-6
    if c0:
        assert dst.to_ncdhw(0, c0).abs().max().item() == 0.0


def merge_case(channels=48, dims=(4, 6, 8), n=2, mode="fp16"):
    torch.manual_seed(3)
    x = torch.randn(n, channels, *dims, device=DEV)
    g, be = torch.randn(8 * channels, device=DEV), torch.randn(8 * channels, device=DEV)
    dst = Blocked(n, 8 * channels, dims[0] // 2, dims[1] // 2, dims[2] // 2, SWIN_MODES[mode], DEV)
    K.swin_merge_ln(_blocked_f32(x), n, channels, *dims, g, be, dst)
    xc = x.double().permute(0, 2, 3, 4, 1)
    cat = torch.cat([xc[:, i::2, j::2, k::2, :] for (i, j, k) in O.MERGE_OFFSETS], -1)
    ref = F.layer_norm(cat, [8 * channels], g.double(), be.double(), 1e-5).permute(0, 4, 1, 2, 3).float()
    err = (dst.to_ncdhw() - ref).abs().max().item()
    print(f"[merge_ln c{channels}] max|err| {err:.2e}")
    assert err < (6e-3 if mode == "fp16" else 5e-2)


def window_attention_case(dims=(8, 9, 10), heads=3, shift=True, n=2, mode="fp16", window=(7, 7, 7)):
    """The attention kernel (pad + roll + partition + bias + mask + softmax + reverse as addressing) vs the oracle's
    explicit pad / roll / window_partition / compute_mask pipeline on the same 16-bit q, k, v."""
    torch.manual_seed(4)
    C_ = heads * 16
    qkv = _rnd16(torch.randn(n, 3 * C_, *dims, device=DEV), mode)
    table = torch.randn((2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1), heads, device=DEV) * 0.5
    qb = _rnd16(torch.randn(3 * C_, device=DEV), mode)
    nm = SWIN_MODES[mode]
    src = Blocked(n, 3 * C_, *dims, nm, DEV)
    K.pack_ncdhw(qkv.contiguous(), src)
    out = Blocked(n, C_, *dims, nm, DEV)
    sh = tuple(w // 2 for w in window) if shift else (0, 0, 0)
    K.swin_window_attention(src, out, table, qb, heads, window, sh)
    got = out.to_ncdhw().cpu()
    # reference on the CPU in fp64
    x = qkv.cpu().double().permute(0, 2, 3, 4, 1)
    d, h, w = dims
    ws, ss = O.get_window_size(dims, window, sh)
    pd, ph, pw = (ws[0] - d % ws[0]) % ws[0], (ws[1] - h % ws[1]) % ws[1], (ws[2] - w % ws[2]) % ws[2]
    xp = qb.cpu().double().view(1, 1, 1, 1, -1).expand(n, d + pd, h + ph, w + pw, 3 * C_).clone()
    xp[:, :d, :h, :w] = x
    dp, hp, wp = xp.shape[1:4]
    mask = None
    if any(s > 0 for s in ss):
        xp = torch.roll(xp, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        mask = O.compute_mask((dp, hp, wp), ws, ss).double()
    xw = O.window_partition(xp, ws)
    b_, nt, _ = xw.shape
    q, k, v = xw.reshape(b_, nt, 3, heads, 16).permute(2, 0, 3, 1, 4)
    att = (q * 0.25) @ k.transpose(-2, -1)
    index = O.relative_position_index(window)
    bias = table.cpu().double()[index[:nt, :nt].reshape(-1)].reshape(nt, nt, -1).permute(2, 0, 1)
    att = att + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        att = (att.view(b_ // nw, nw, heads, nt, nt) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, nt, nt)
    o = (att.softmax(-1) @ v).transpose(1, 2).reshape(b_, nt, C_)
    o = O.window_reverse(o.view(-1, ws[0], ws[1], ws[2], C_), ws, (n, dp, hp, wp))
    if any(s > 0 for s in ss):
        o = torch.roll(o, shifts=ss, dims=(1, 2, 3))
    ref = o[:, :d, :h, :w].permute(0, 4, 1, 2, 3).float()
    err = (got - ref).abs().max().item()
    tol = 6e-3 if mode == "fp16" else 4e-2
    print(f"[window_attention {dims} heads {heads} shift {sh}] max|err| {err:.2e} (ref max {ref.abs().max().item():.2f}, tol {tol})")
    assert err < tol


def resnorm_case(channels=48, dims=(6, 7, 8), n=2, mode="fp16", identity=False):
    torch.manual_seed(5)
    nm = SWIN_MODES[mode]
    a = torch.randn(n, channels, *dims, device=DEV) * 3 + 1
    r = torch.randn(n, channels, *dims, device=DEV) * 2 - 1
    vox = dims[0] * dims[1] * dims[2]

    def table(t):
        m = t.double().mean((2, 3, 4))
        v = t.double().var((2, 3, 4), unbiased=False)
        return torch.stack([m, 1.0 / torch.sqrt(v + 1e-5)], -1).float().contiguous()

    dst = Blocked(n, channels, *dims, nm, DEV)
    if identity:
        rb = Blocked(n, 2 * channels, *dims, nm, DEV)
        rb.t.zero_()
        K.pack_ncdhw(r.contiguous(), rb, channels)
        r16 = rb.to_ncdhw(channels, channels)
        K.instnorm_residual_act(_blocked_f32(a), True, table(a), rb.t, False, None, rb.cbt, channels, dst, 0, n, channels, vox, 0.01)
        ref = F.leaky_relu(F.instance_norm(a.double(), eps=1e-5) + r16.double(), 0.01).float()
    else:
        K.instnorm_residual_act(_blocked_f32(a), True, table(a), _blocked_f32(r), True, table(r), channels // 8, 0, dst, 0, n,
                                channels, vox, 0.01)
        ref = F.leaky_relu(F.instance_norm(a.double(), eps=1e-5) + F.instance_norm(r.double(), eps=1e-5), 0.01).float()
    err = (dst.to_ncdhw() - ref).abs().max().item()
    print(f"[resnorm identity{int(identity)}] max|err| {err:.2e}")
    assert err < (6e-3 if mode == "fp16" else 5e-2)


def swin_unetr_case(size=64, n=1, mode="fp16", feature_size=48, check_hidden=True):
    """Whole SwinUNETR forward vs the oracle (fp32 CPU).  fp16 operands: gates as for the UNet's fp16 rung."""
    from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
    torch.manual_seed(0)
    m = SwinUNETR(in_channels=2, out_channels=8, feature_size=feature_size).eval()
    with torch.no_grad():   # make the position bias and the LayerNorm affines non-trivial
        for name, p in m.named_parameters():
            if "relative_position_bias_table" in name:
                p.normal_(0, 0.5)
            elif "norm" in name and name.endswith("weight"):
                p.uniform_(0.5, 1.5)
            elif "norm" in name and name.endswith("bias"):
                p.normal_(0, 0.2)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(n, 2, size, size, size)
    ref, hs = O.swin_unetr_forward(sd, x, return_hidden=True)
    m = m.cuda().set_numeric_mode(mode)
    with torch.no_grad():
        got, ghs = m(x.cuda(), return_features=True)
        got2 = m(x.cuda())
    assert torch.equal(got, got2), "two forwards differ"
    if check_hidden:
        for i, (a, b) in enumerate(zip(ghs, hs)):
            e = (a.cpu() - b).norm() / b.norm()
            print(f"[swin_unetr {size}^3 {mode}] hidden[{i}] {tuple(b.shape)} rel-L2 {e.item():.2e}")
            assert e.item() < (2e-2 if mode == "fp16" else 8e-2)
    d = got.cpu() - ref
    max_abs, rel = d.abs().max().item(), (d.norm() / ref.norm()).item()
    agree = (got.cpu().argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"[swin_unetr {size}^3 n{n} {mode}] logits max|err| {max_abs:.2e} rel-L2 {rel:.2e} labels {agree * 100:.3f}% "
          f"(ref std {ref.std().item():.3f})")
    if mode == "fp16":
        assert rel < 1e-2 and agree > 0.99
    else:
        assert rel < 6e-2 and agree > 0.95
    return max_abs, rel, agree


def swin_sliding_window_case(vol_shape=(64, 96, 96), roi=64, overlap=0.25, blend="gaussian"):
    """SwinUNETR under the sliding-window inferer (the reference's Trainer.predict path, trainer.py:370-395): windows
    gathered by kernels (blocked copy for encoder1 + fp32 batch for the patch embedding), batched engine forward in a
    CUDA graph, the 1x1x1 head fused into the blend — vs the oracle's MONAI-style loop around the oracle model."""
    from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
    from mmseg_b200.src.trainer.inference import sliding_window_inference
    from oracle.sliding_window import sliding_window_inference as oracle_swi
    torch.manual_seed(0)
    m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    vol = torch.randn(1, 2, *vol_shape)
    want = oracle_swi(vol, (roi,) * 3, 4, lambda w: O.swin_unetr_forward(sd, w), overlap=overlap, mode=blend)
    m = m.cuda()
    got = sliding_window_inference(vol.cuda(), (roi,) * 3, 4, m, overlap=overlap, mode=blend).cpu()
    got2 = sliding_window_inference(vol.cuda(), (roi,) * 3, 4, m, overlap=overlap, mode=blend).cpu()   # graph replay
    assert torch.equal(got, got2)
    d = got - want
    rel = (d.norm() / want.norm()).item()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    print(f"[swin sliding window {vol_shape} roi {roi}] max|err| {d.abs().max().item():.2e} rel-L2 {rel:.2e} labels {agree * 100:.3f}%")
    assert rel < 5e-3 and agree > 0.995
