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
    shape = (size, size, size) if isinstance(size, int) else tuple(size)
    x = torch.randn(n, 2, *shape)
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


# ------------------------------------------------------------------------------------------------ training (backward)
def _b16(x):
    """NCDHW fp32 (values already bf16-representable) -> blocked bf16 tensor."""
    n, c, Z, Y, X = x.shape
    return x.view(n, c // 8, 8, Z, Y, X).permute(0, 1, 3, 4, 5, 2).contiguous().to(torch.bfloat16)


def _u16(xb):
    n, cb, Z, Y, X, _ = xb.shape
    return xb.float().permute(0, 1, 5, 2, 3, 4).reshape(n, cb * 8, Z, Y, X)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _bf(x):
    return x.to(torch.bfloat16).float()


def ln_backward_case(channels=96, dims=(4, 5, 6), n=2, with_add=True):
    """AddLayerNormFn / LayerNormFn (kernels) vs torch autograd in fp64 on the same inputs."""
    from mmseg_b200 import swin_train as T
    torch.manual_seed(7)
    x = torch.randn(n, channels, *dims, device=DEV) * 2 + 0.3
    y = _bf(torch.randn(n, channels, *dims, device=DEV))
    g = (torch.rand(channels, device=DEV) + 0.5).requires_grad_(True)
    be = torch.randn(channels, device=DEV).requires_grad_(True)
    go_x = torch.randn(n, channels, *dims, device=DEV)
    go_ln = _bf(torch.randn(n, channels, *dims, device=DEV))
    xs = _blocked_f32(x).requires_grad_(True)
    y16 = _b16(y).requires_grad_(True)
    if with_add:
        xs_out, ln = T.AddLayerNormFn.apply(xs, y16, g, be, 1e-5)
        torch.autograd.backward([xs_out, ln], [_blocked_f32(go_x), _b16(go_ln)])
    else:
        ln = T.LayerNormFn.apply(xs, g, be, 1e-5)
        ln.backward(_b16(go_ln))
    # reference
    xr = x.double().requires_grad_(True)
    yr = y.double().requires_grad_(True)
    gr, br = g.detach().double().requires_grad_(True), be.detach().double().requires_grad_(True)
    s = xr + yr if with_add else xr
    lnr = F.layer_norm(s.permute(0, 2, 3, 4, 1), [channels], gr, br, 1e-5).permute(0, 4, 1, 2, 3)
    loss = (lnr * go_ln.double()).sum() + ((s * go_x.double()).sum() if with_add else 0.0)
    loss.backward()
    e = {"ln": _rel(_u16(ln.detach()), lnr.detach()), "dx": _rel(_unblocked_f32(xs.grad), xr.grad),
         "dgamma": _rel(g.grad, gr.grad), "dbeta": _rel(be.grad, br.grad)}
    if with_add:
        e["dy"] = _rel(_u16(y16.grad), yr.grad)
    print(f"[ln backward c{channels} add{int(with_add)}] " + " ".join(f"{k} {v:.2e}" for k, v in e.items()))
    assert e["ln"] < 5e-3 and e["dx"] < 1e-4 and e["dgamma"] < 1e-4 and e["dbeta"] < 1e-5 and e.get("dy", 0) < 5e-3


def gelu_merge_patch_backward_case():
    from mmseg_b200 import swin_train as T
    torch.manual_seed(8)
    # GELU
    h = _bf(torch.randn(2, 32, 3, 4, 5, device=DEV) * 2)
    go = _bf(torch.randn_like(h))
    h16 = _b16(h).requires_grad_(True)
    ident = torch.zeros((2, 32, 2), device=DEV)
    ident[:, :, 1] = 1
    T.GeluFn.apply(h16, ident).backward(_b16(go))
    hr = h.double().requires_grad_(True)
    (F.gelu(hr) * go.double()).sum().backward()
    e_gelu = _rel(_u16(h16.grad), hr.grad)
    # merge gather / scatter
    x = torch.randn(2, 16, 4, 6, 8, device=DEV)
    xs = _blocked_f32(x).requires_grad_(True)
    cat = T.MergeGatherFn.apply(xs)
    gcat = torch.randn_like(cat)
    cat.backward(gcat)
    xr = x.double().requires_grad_(True)
    xc = xr.permute(0, 2, 3, 4, 1)
    catr = torch.cat([xc[:, i::2, j::2, k::2, :] for (i, j, k) in O.MERGE_OFFSETS], -1).permute(0, 4, 1, 2, 3)
    e_cat = (_unblocked_f32(cat.detach()) - catr.detach().float()).abs().max().item()
    (catr * _unblocked_f32(gcat).double()).sum().backward()
    e_scat = (_unblocked_f32(xs.grad) - xr.grad.float()).abs().max().item()
    # patch embedding weight / bias gradient
    img = torch.randn(2, 2, 12, 8, 16, device=DEV)
    w = (torch.randn(48, 2, 2, 2, 2, device=DEV) * 0.3).requires_grad_(True)
    b = torch.randn(48, device=DEV).requires_grad_(True)
    out = T.PatchEmbedFn.apply(img, w, b)
    gout = torch.randn_like(out)
    out.backward(gout)
    wr, br = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    (F.conv3d(img.double(), wr, br, stride=2) * _unblocked_f32(gout).double()).sum().backward()
    e_pw, e_pb = _rel(w.grad, wr.grad), _rel(b.grad, br.grad)
    print(f"[gelu / merge / patch-embed backward] gelu {e_gelu:.2e} gather {e_cat:.1e} scatter {e_scat:.1e} dW {e_pw:.1e} db {e_pb:.1e}")
    assert e_gelu < 6e-3 and e_cat == 0.0 and e_scat < 1e-6 and e_pw < 1e-5 and e_pb < 1e-5


def window_attention_backward_case(dims=(8, 9, 10), heads=2, shift=True, n=1, window=(7, 7, 7)):
    """WindowAttentionFn backward (dqkv, bias-table gradient, qkv-bias gradient of the padded tokens) vs fp64 autograd of
    the oracle's explicit pad / roll / partition pipeline on the same bf16 q, k, v."""
    from mmseg_b200 import swin_train as T
    torch.manual_seed(9)
    C_ = heads * 16
    qkv = _bf(torch.randn(n, 3 * C_, *dims, device=DEV))
    table = (torch.randn((2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1), heads, device=DEV) * 0.5).requires_grad_(True)
    qb = _bf(torch.randn(3 * C_, device=DEV)).requires_grad_(True)
    go = _bf(torch.randn(n, C_, *dims, device=DEV))
    sh = tuple(w // 2 for w in window) if shift else (0, 0, 0)
    q16 = _b16(qkv).requires_grad_(True)
    out = T.WindowAttentionFn.apply(q16, table, qb, heads, window, sh)
    out.backward(_b16(go))
    # reference (CPU fp64)
    xr = qkv.cpu().double().requires_grad_(True)
    tr = table.detach().cpu().double().requires_grad_(True)
    br = qb.detach().cpu().double().requires_grad_(True)
    x = xr.permute(0, 2, 3, 4, 1)
    d, h, w = dims
    ws, ss = O.get_window_size(dims, window, sh)
    pd, ph, pw = (ws[0] - d % ws[0]) % ws[0], (ws[1] - h % ws[1]) % ws[1], (ws[2] - w % ws[2]) % ws[2]
    inside = torch.zeros(n, d + pd, h + ph, w + pw, 1, dtype=torch.float64)
    inside[:, :d, :h, :w] = 1
    xp = F.pad(x, (0, 0, 0, pw, 0, ph, 0, pd)) + (1 - inside) * br.view(1, 1, 1, 1, -1)
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
    att = att + tr[index[:nt, :nt].reshape(-1)].reshape(nt, nt, -1).permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        att = (att.view(b_ // nw, nw, heads, nt, nt) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, nt, nt)
    o = (att.softmax(-1) @ v).transpose(1, 2).reshape(b_, nt, C_)
    o = O.window_reverse(o.view(-1, ws[0], ws[1], ws[2], C_), ws, (n, dp, hp, wp))
    if any(s > 0 for s in ss):
        o = torch.roll(o, shifts=ss, dims=(1, 2, 3))
    ref = o[:, :d, :h, :w].permute(0, 4, 1, 2, 3)
    (ref * go.cpu().double()).sum().backward()
    e = {"out": _rel(_u16(out.detach()).cpu(), ref.detach()), "dqkv": _rel(_u16(q16.grad).cpu(), xr.grad),
         "dtable": _rel(table.grad.cpu(), tr.grad), "dqkv_bias": _rel(qb.grad.cpu(), br.grad)}
    print(f"[window attention backward {dims} heads {heads} shift {sh}] " + " ".join(f"{k_} {v_:.2e}" for k_, v_ in e.items()))
    assert e["out"] < 6e-3 and e["dqkv"] < 1e-2 and e["dtable"] < 5e-3
    if br.grad.abs().max() > 0:
        assert e["dqkv_bias"] < 1e-2


def res_block_backward_case(cin=32, cout=48, dims=(6, 8, 10), n=2):
    """UnetResBlock through ConvStatsFn / NormActFn / ResTailFn vs fp64 autograd (bf16-rounded weights and input)."""
    from mmseg_b200 import swin_train as T
    from mmseg_b200.src.models.backbones.swin_unetr import UnetResBlock
    torch.manual_seed(10)
    blk = UnetResBlock(cin, cout).to(DEV)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(_bf(p))
    x = _bf(torch.randn(n, cin, *dims, device=DEV))
    go = _bf(torch.randn(n, cout, *dims, device=DEV))
    x16 = _b16(x).requires_grad_(True)
    y = T._res_block(blk, x16, [(0, cin)])
    y.backward(_b16(go))
    xr = x.double().requires_grad_(True)
    ws = {k: v.detach().double().requires_grad_(True) for k, v in blk.named_parameters()}
    o = F.leaky_relu(F.instance_norm(F.conv3d(xr, ws["conv1.conv.weight"], padding=1)), 0.01)
    o = F.instance_norm(F.conv3d(o, ws["conv2.conv.weight"], padding=1))
    r = F.instance_norm(F.conv3d(xr, ws["conv3.conv.weight"])) if cin != cout else xr
    yr = F.leaky_relu(o + r, 0.01)
    (yr * go.double()).sum().backward()
    e = {"y": _rel(_u16(y.detach()), yr.detach()), "dx": _rel(_u16(x16.grad), xr.grad)}
    for k, p in blk.named_parameters():
        e["d" + k.split(".")[0]] = _rel(p.grad, ws[k].grad)
    print(f"[res block backward {cin}->{cout}] " + " ".join(f"{k} {v:.2e}" for k, v in e.items()))
    # ~4e-2 is the bf16 floor of this block: activations within bf16 noise of 0 take the other LeakyReLU slope (a fraction f of
    # flipped elements shows up as sqrt(f) in relative L2); every kernel on the way is checked tightly on its own
    assert e["y"] < 1e-2 and all(v < 8e-2 for v in e.values())


def swin_train_step_case(size=64, n=1):
    """Whole-model gradient check: DiceCE loss and every parameter gradient of SwinUNETR vs fp64 autograd of the oracle."""
    from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
    from mmseg_b200.src.trainer.losses import DiceCELoss
    import oracle.losses as OL
    torch.manual_seed(0)
    m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).train()
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "relative_position_bias_table" in name:
                p.normal_(0, 0.3)
    sd64 = {k: v.detach().double().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    shape = (size, size, size) if isinstance(size, int) else tuple(size)
    x = torch.randn(n, 2, *shape)
    y = torch.randint(0, 8, (n, *shape))
    # oracle in fp64 with autograd (the oracle detaches its parameters: bypass that here)
    keep = O._p
    O._p = lambda sd, key, dtype: sd[key]
    try:
        ref_logits = O.swin_unetr_forward(sd64, x, dtype=torch.float64)
        # DiceCE as oracle/losses.py states it (reference losses.py:216-228), kept differentiable here
        p_ = torch.softmax(ref_logits, dim=1).flatten(2)
        t_ = F.one_hot(y, 8).movedim(-1, 1).double().flatten(2)
        dice = (1.0 - (2.0 * (p_ * t_).sum(-1) + 1.0) / (p_.sum(-1) + t_.sum(-1) + 1.0)).mean()
        ref_loss = 0.5 * dice + 0.5 * F.cross_entropy(ref_logits, y)
        assert abs(ref_loss.item() - OL.dice_ce_loss(ref_logits, y, dtype=torch.float64)[0].item()) < 1e-9
        ref_loss.backward()
    finally:
        O._p = keep
    m = m.cuda()
    loss_fn = DiceCELoss()
    logits = m(x.cuda())
    loss = loss_fn(logits, y.cuda())
    loss.backward()
    rl = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
    worst, rows = 0.0, []
    for name, p in m.named_parameters():
        g_ref = sd64[name].grad
        assert p.grad is not None, name
        e = _rel(p.grad.cpu(), g_ref)
        rows.append((e, name))
        worst = max(worst, e)
    rows.sort(reverse=True)
    med = rows[len(rows) // 2][0]
    print(f"[swin train step {size}^3] loss {loss.item():.5f} vs oracle {ref_loss.item():.5f} (rel {rl:.1e}); parameter-gradient "
          f"rel-L2: median {med:.2e}, worst {worst:.2e} ({rows[0][1]}); next: " + ", ".join(f"{n_}: {e:.1e}" for e, n_ in rows[1:4]))
    # The error grows smoothly with depth (decoder1 2-4 %, ... deepest transformer parameters 22-33 %): the bf16 floor of a
    # 60-layer path with LeakyReLU / GELU decisions — the reference's own arithmetic under torch.autocast(bfloat16) shows
    # median 0.22 / worst 0.35 against fp64 on this configuration (tools/swin_train_check.py), the kernels 0.20 / 0.33.
    # Every backward kernel is checked tightly on its own above; here: the loss, the bound and the gradient norms.
    norm_ratio = [(p.grad.norm().item() / sd64[name].grad.norm().item()) for name, p in m.named_parameters()]
    assert rl < 1e-3 and med < 0.3 and worst < 0.45 and 0.8 < min(norm_ratio) and max(norm_ratio) < 1.25


def swin_trainer_case(tmp_dir):
    """The reference-facing Trainer with model.name = swin_unetr (main.py --mode train / inference): train() through the
    kernel autograd path, validation, predict_array() through the sliding-window engine."""
    import numpy as np
    from mmseg_b200.src.models.build import build_model
    from mmseg_b200.src.trainer import Trainer
    torch.manual_seed(0)
    cfg = {"model": {"name": "swin_unetr", "in_channels": 2, "out_channels": 4,
                     "backbone": {"feature_size": 48, "img_size": [32, 32, 32]}, "fusion": {"type": "early"},
                     "head": {"dropout": 0.0}},
           "data": {"modalities": ["CT", "PET"]},
           "hardware": {"device": "cuda", "mixed_precision": True},
           "training": {"epochs": 2, "accumulation_steps": 1,
                        "optimizer": {"name": "adamw", "lr": 1e-3, "weight_decay": 1e-5},
                        "scheduler": {"name": "cosine"}, "loss": {"name": "dice_ce"},
                        "checkpoint": {"save_last": True, "save_best": False}},
           "inference": {"batch_size": 2, "sliding_window": {"roi_size": [64, 64, 64], "overlap": 0.25}},
           "experiment": {"output_dir": str(tmp_dir), "name": "swin"}}
    g = torch.Generator().manual_seed(5)

    def batches(n):
        out = []
        for _ in range(n):
            lab = torch.randint(0, 4, (1, 8, 8, 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3)
            img = torch.randn(1, 2, 64, 64, 64, generator=g) * 0.3 + lab[:, None].float()
            out.append({"image": img, "label": lab})
        return out
    tr = Trainer(cfg, build_model(cfg), train_loader=batches(6), val_loader=batches(1))
    hist = tr.train()
    assert all(np.isfinite(v) for v in hist["train_loss"] + hist["val_loss"])
    assert hist["train_loss"][1] < hist["train_loss"][0], hist
    vol = batches(1)[0]["image"][0].numpy().astype(np.float32)[:, :, :, :64]
    pred = tr.predict_array(vol)
    assert pred.shape == (64, 64, 64) and pred.dtype == np.uint8
    with torch.no_grad():
        direct = tr.model(torch.from_numpy(vol)[None].cuda()).argmax(1)[0].cpu().numpy()
    agree = float((pred == direct).mean())
    print(f"[swin trainer] losses {hist['train_loss']} val dice {hist['val_dice']}; predict_array vs direct forward labels {agree * 100:.3f}%")
    assert agree > 0.995
