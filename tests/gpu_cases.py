"""GPU parity cases: CUDA path (through the C ABI) vs the oracle / torch fp32 on the same seeded inputs.

Plain functions (assert-based) so that both pytest (`tests/test_gpu_*.py`, marker `gpu`) and the stand-alone runner
`tools/gpu_check.py` (one subprocess per case, so a trapped kernel cannot poison the next case) can call them.
"""
import types

import torch
import torch.nn.functional as F

import mmseg_b200  # noqa: F401
from mmseg_b200 import _lib
from mmseg_b200 import kernels as K
from mmseg_b200.engine import ConvRunner
from mmseg_b200.kernels import Blocked
from mmseg_b200.tiling import plan_conv, ConvTile

DEV = "cuda"


def _bf(x):
    return x.to(torch.bfloat16).float()


def _ref_conv(x, w, b=None, pad=1):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return F.conv3d(x.double(), w.double(), None if b is None else b.double(), padding=pad).float()


def _report(name, got, ref, tol):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"[{name}] max|err|={err:.3e} (ref max {scale:.3e}, tol {tol:.1e})", flush=True)
    return err


def _make_tile(pw, TX, TY, TZ, ks, stages=3):
    h = ks // 2
    mt = ((TY - 1) * (TX + 2 * h) + TX + 127) // 128
    return ConvTile(TX, TY, TZ, pw.NT, pw.n_ntiles, stages, mt, 0, 0, 0.0)


def conv_case(cin, cout, shape, n_img=1, tile=None, split=False, flags=0, ks=3, seed=0, name="conv"):
    """raw conv output (blocked) + InstanceNorm partial statistics vs F.conv3d in fp64."""
    torch.manual_seed(seed)
    Z, Y, X = shape
    x = torch.randn(n_img, cin, Z, Y, X, device=DEV)
    w = torch.randn(cout, cin, ks, ks, ks, device=DEV) * (1.0 / (cin * ks ** 3) ** 0.5)
    if not split:
        x, w = _bf(x), _bf(w)
    pw = K.pack_conv_weight(w, None, split, [cin], use_bias=False)
    src = Blocked(n_img, (cin + 15) // 16 * 16, Z, Y, X, split, DEV)
    K.pack_ncdhw(x, src)
    a_cb = K.a_chunk_table(src, [0], [cin], split)
    t = plan_conv(X, Y, Z, n_img, pw.n_kchunks, pw.n_out, ks, pw.NT) if tile is None else _make_tile(pw, *tile, ks)
    raw = torch.full((n_img, pw.n_out // 8, Z, Y, X, 8), float("nan"), device=DEV,
                     dtype=torch.float32 if split else torch.bfloat16)
    tiles_per_img = ((X + t.TX - 1) // t.TX) * ((Y + t.TY - 1) // t.TY) * ((Z + t.TZ - 1) // t.TZ)
    stats = torch.zeros((n_img, tiles_per_img, pw.n_out, 2), device=DEV)
    K.conv3d(src, pw, a_cb, raw, _lib.OUT_BLOCKED_F32 if split else _lib.OUT_BLOCKED_BF16, stats=stats,
             dst_cbt=pw.n_out // 8, tile=t, flags=flags)
    torch.cuda.synchronize()
    got = raw.float().permute(0, 1, 5, 2, 3, 4).reshape(n_img, pw.n_out, Z, Y, X)[:, :cout]
    ref = _ref_conv(x, w, pad=ks // 2)
    tol = 3e-5 if split else 2e-2
    err = _report(f"{name} cin={cin} cout={cout} {shape} n={n_img} tile={tile} split={split} flags={flags}", got, ref, tol)
    assert torch.isfinite(got).all(), "non-finite output (unwritten voxels?)"
    assert err <= tol * max(1.0, ref.abs().max().item())
    # statistics: sum and sum of squares over voxels per (img, channel)
    s = stats.sum(1)[:, :cout]
    ref64 = F.conv3d(x.double(), w.double(), padding=ks // 2)
    s1, s2 = ref64.sum(dim=(2, 3, 4)), (ref64 ** 2).sum(dim=(2, 3, 4))
    e1 = (s[..., 0].double() - s1).abs().max().item() / max(1.0, s2.max().item() ** 0.5)
    e2 = ((s[..., 1].double() - s2).abs() / s2).max().item()
    print(f"   stats: sum err {e1:.2e}, sumsq rel err {e2:.2e}", flush=True)
    assert e1 < 1e-2 and e2 < 1e-3
    return err


def conv_block_case(cin, cout, shape, n_img=1, split=False, pooled=True, slope=0.0):
    """conv -> IN -> act (+pool) twice through ConvRunner vs torch fp64."""
    torch.manual_seed(3)
    Z, Y, X = shape
    x = torch.randn(n_img, cin, Z, Y, X, device=DEV)
    w1 = torch.randn(cout, cin, 3, 3, 3, device=DEV) * (1.0 / (cin * 27) ** 0.5)
    w2 = torch.randn(cout, cout, 3, 3, 3, device=DEV) * (1.0 / (cout * 27) ** 0.5)
    r = ConvRunner(split, DEV)
    p1 = K.pack_conv_weight(w1, None, split, [cin], use_bias=False)
    p2 = K.pack_conv_weight(w2, None, split, None, use_bias=False)
    src = Blocked(n_img, (cin + 15) // 16 * 16, Z, Y, X, split, DEV)
    K.pack_ncdhw(x, src)
    mid = Blocked(n_img, cout, Z, Y, X, split, DEV)
    out = Blocked(n_img, cout, Z, Y, X, split, DEV)
    pl = Blocked(n_img, cout, Z // 2, Y // 2, X // 2, split, DEV) if pooled else None
    r.conv_norm_act(src, [(0, cin)], p1, mid, slope=slope)
    r.conv_norm_act(mid, [(0, cout)], p2, out, pooled=pl, slope=slope)
    torch.cuda.synchronize()
    act = (lambda t: F.leaky_relu(t, slope)) if slope else F.relu
    xd = x.double()
    y = act(F.instance_norm(F.conv3d(xd, w1.double(), padding=1), eps=1e-5))
    y = act(F.instance_norm(F.conv3d(y, w2.double(), padding=1), eps=1e-5)).float()
    tol = 2e-4 if split else 8e-2
    err = _report(f"conv_block cin={cin} cout={cout} {shape} split={split} slope={slope}", out.to_ncdhw(), y, tol)
    assert err <= tol
    if pooled:
        e2 = _report("   pooled", pl.to_ncdhw(), F.max_pool3d(y, 2), tol)
        assert e2 <= tol
    return err


def convt_case(cin, shape, n_img=1, split=False):
    torch.manual_seed(4)
    Z, Y, X = shape
    cout = cin // 2
    x = torch.randn(n_img, cin, Z, Y, X, device=DEV)
    w = torch.randn(cin, cout, 2, 2, 2, device=DEV) * (1.0 / cin ** 0.5)
    b = torch.randn(cout, device=DEV)
    if not split:
        x, w = _bf(x), _bf(w)
    r = ConvRunner(split, DEV)
    pw = K.pack_conv_weight(w, b, split, None, transposed=True)
    src = Blocked(n_img, cin, Z, Y, X, split, DEV)
    K.pack_ncdhw(x, src)
    cat = Blocked(n_img, 2 * cout, 2 * Z, 2 * Y, 2 * X, split, DEV)
    cat.t.zero_()
    r.conv_transpose(src, [(0, cin)], pw, cat, 0)
    torch.cuda.synchronize()
    ref = F.conv_transpose3d(x.double(), w.double(), b.double(), stride=2).float()
    tol = 3e-5 if split else 2e-2
    err = _report(f"convT cin={cin} {shape} split={split}", cat.to_ncdhw(0, cout), ref, tol)
    assert err <= tol * max(1.0, ref.abs().max().item())
    assert cat.to_ncdhw(cout, cout).abs().max().item() == 0.0, "convT wrote outside its half of the concat buffer"
    return err


def logits_case(cin, k, shape, n_img=2, split=False):
    torch.manual_seed(5)
    Z, Y, X = shape
    x = torch.randn(n_img, cin, Z, Y, X, device=DEV)
    w = torch.randn(k, cin, 1, 1, 1, device=DEV) * (1.0 / cin ** 0.5)
    b = torch.randn(k, device=DEV)
    if not split:
        x, w = _bf(x), _bf(w)
    r = ConvRunner(split, DEV)
    pw = K.pack_conv_weight(w, b, split, None)
    src = Blocked(n_img, cin, Z, Y, X, split, DEV)
    K.pack_ncdhw(x, src)
    out = torch.full((n_img, k, Z, Y, X), float("nan"), device=DEV)
    r.conv_logits(src, [(0, cin)], pw, out)
    torch.cuda.synchronize()
    ref = F.conv3d(x.double(), w.double(), b.double()).float()
    tol = 3e-5 if split else 2e-2
    err = _report(f"logits 1x1 cin={cin} k={k} {shape} split={split}", out, ref, tol)
    assert err <= tol * max(1.0, ref.abs().max().item())
    return err


def pack_roundtrip_case():
    torch.manual_seed(6)
    x = torch.randn(2, 5, 4, 6, 10, device=DEV)
    for split in (False, True):
        b = Blocked(2, 16, 4, 6, 10, split, DEV)
        K.pack_ncdhw(x, b)
        y = b.to_ncdhw(0, 5)
        tol = 1e-5 if split else 1e-2
        err = _report(f"pack/unpack split={split}", y, x, tol)
        assert err <= tol * 4
        assert b.to_ncdhw(8, 8).abs().max().item() == 0.0


def _metrics(got, ref):
    d = (got - ref).double()
    max_abs = d.abs().max().item()
    rel_l2 = (d.norm() / ref.double().norm()).item()
    agree = (got.argmax(1) == ref.argmax(1)).double().mean().item()
    return max_abs, rel_l2, agree


def unet_case(features=(16, 32, 64), S=32, n_img=1, mode="parity", in_ch=2, seed=0, tol=None):
    """Drop-in UNet3D (CUDA kernels) vs the oracle restatement (CPU fp32) on the same state_dict and input."""
    from mmseg_b200.src.models.backbones.unet import UNet3D
    from oracle.models import unet3d_forward
    torch.manual_seed(seed)
    m = UNet3D(in_channels=in_ch, out_channels=8, features=list(features)).eval()
    sd = {"backbone." + k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(n_img, in_ch, S, S, S)
    ref = unet3d_forward(sd, x)
    m = m.to(DEV).set_numeric_mode(mode)
    with torch.no_grad():
        got = m(x.to(DEV)).cpu()
    max_abs, rel_l2, agree = _metrics(got, ref)
    print(f"[unet {features} S={S} n={n_img} mode={mode}] max_abs={max_abs:.3e} rel_l2={rel_l2:.3e} "
          f"label_agree={agree * 100:.4f}%", flush=True)
    if tol is None:
        tol = (2e-2, 1e-3, 0.999) if mode == "parity" else (2e-1, 5e-2, 0.95)
    assert torch.isfinite(got).all()
    assert max_abs <= tol[0] and rel_l2 <= tol[1] and agree >= tol[2], (max_abs, rel_l2, agree)
    return max_abs, rel_l2, agree


def unet_time_case(n_img=1, S=96, mode="bf16", iters=5):
    from mmseg_b200.src.models.backbones.unet import UNet3D
    torch.manual_seed(0)
    m = UNet3D(in_channels=2, out_channels=8).eval().to(DEV).set_numeric_mode(mode)
    x = torch.randn(n_img, 2, S, S, S, device=DEV)
    with torch.no_grad():
        for _ in range(2):
            y = m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gf = 403.2 * n_img * (S / 96) ** 3
    print(f"[unet time n={n_img} S={S} mode={mode}] {ms:.3f} ms/forward -> {gf / ms:.1f} TFLOP/s algorithmic, "
          f"{n_img * S ** 3 / ms / 1e3:.2f} Mvox/s", flush=True)
    assert torch.isfinite(y).all()
    return ms
