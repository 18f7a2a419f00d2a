"""Host logic of the conv path (packing, K-chunk tables, planner) checked by emulating the kernel's data path."""
import types

import pytest
import torch
import torch.nn.functional as F

import mmseg_b200  # noqa: F401
from mmseg_b200 import kernels as K
from mmseg_b200.tiling import plan_conv, smem_bytes, ConvTile
from tests.emulate import emulate_conv, to_blocked


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("cin,cout,shape,tile", [
    (2, 16, (5, 6, 7), None),
    (16, 32, (4, 9, 10), (10, 4, 2)),
    (32, 16, (3, 5, 20), (10, 3, 2)),      # ragged x/y/z tiles
])
def test_conv3x3_emulated(cin, cout, shape, tile):
    torch.manual_seed(0)
    Z, Y, X = shape
    x = _bf(torch.randn(2, cin, Z, Y, X))
    w = _bf(torch.randn(cout, cin, 3, 3, 3) * 0.2)
    pw = K.pack_conv_weight(w, None, False, [cin], use_bias=False)
    stub = types.SimpleNamespace(lo_off=0)
    a_cb = K.a_chunk_table(stub, [0], [cin], False)
    if tile is None:
        t = plan_conv(X, Y, Z, 2, pw.n_kchunks, pw.n_out, 3, pw.NT)
    else:
        TX, TY, TZ = tile
        mt = ((TY - 1) * (TX + 2) + TX + 127) // 128
        t = ConvTile(TX, TY, TZ, pw.NT, pw.n_ntiles, 2, mt, 0, 0, 0.0)
    src = to_blocked(x)
    got = emulate_conv(src, src.shape[1], pw, a_cb, t, 2, Z, Y, X)[:, :cout]
    ref = F.conv3d(x, w, padding=1)
    assert torch.allclose(got, ref, atol=1e-3, rtol=1e-4), (got - ref).abs().max()


def test_concat_segments_and_split_mode():
    """Two-segment K (cat([up, skip])) with hi/lo split operands reproduces the fp32 conv to ~2^-16."""
    torch.manual_seed(1)
    Z, Y, X = 3, 4, 6
    up, skip = torch.randn(1, 16, Z, Y, X), torch.randn(1, 16, Z, Y, X)
    w = torch.randn(16, 32, 3, 3, 3) * 0.1
    pw = K.pack_conv_weight(w, None, True, [16, 16], use_bias=False)
    assert pw.n_kchunks == 6
    # buffer = [hi(up) hi(skip) | lo(up) lo(skip)], cb = 4 per plane
    cat = torch.cat([up, skip], 1)
    hi = _bf(cat)
    lo = _bf(cat - hi)
    src = torch.cat([to_blocked(hi), to_blocked(lo)], dim=1)
    stub = types.SimpleNamespace(lo_off=4)
    a_cb = K.a_chunk_table(stub, [0, 16], [16, 16], True)
    assert a_cb == [0, 2, 4, 6, 0, 2]
    t = plan_conv(X, Y, Z, 1, pw.n_kchunks, pw.n_out, 3, pw.NT)
    got = emulate_conv(src, 8, pw, a_cb, t, 1, Z, Y, X)[:, :16]
    ref = F.conv3d(cat, w, padding=1)
    assert (got - ref).abs().max() < 2e-4 * ref.abs().max()


def test_conv_transpose_as_gemm():
    torch.manual_seed(2)
    Z, Y, X = 2, 3, 4
    x = _bf(torch.randn(1, 32, Z, Y, X))
    w = _bf(torch.randn(32, 16, 2, 2, 2) * 0.2)
    b = torch.randn(16)
    pw = K.pack_conv_weight(w, b, False, None, transposed=True)
    assert pw.n_out == 128 and pw.ksize == 1
    stub = types.SimpleNamespace(lo_off=0)
    a_cb = K.a_chunk_table(stub, [0], [32], False)
    t = plan_conv(X, Y, Z, 1, pw.n_kchunks, pw.n_out, 1, pw.NT)
    g = emulate_conv(to_blocked(x), 4, pw, a_cb, t, 1, Z, Y, X)  # [1, 8*16, Z, Y, X]
    g = g + pw.bias.view(1, -1, 1, 1, 1)
    out = torch.zeros(1, 16, 2 * Z, 2 * Y, 2 * X)
    CB = 2  # 16 output channels = 2 blocks; column n = (((dz*2+dy)*CB + cb)*2 + dx)*8 + j
    for n in range(128):
        j, dx, cb, tzy = n % 8, (n // 8) % 2, (n // 16) % CB, n // (16 * CB)
        out[:, cb * 8 + j, (tzy >> 1)::2, (tzy & 1)::2, dx::2] = g[:, n]
    ref = F.conv_transpose3d(x, w, b, stride=2)
    assert torch.allclose(out, ref, atol=1e-3, rtol=1e-4)


def test_planner_matches_library_limits():
    """Every plan the Python planner emits must be accepted by the C-side plan_conv (same smem arithmetic)."""
    import ctypes as C
    from mmseg_b200 import _lib
    for (X, Y, Z, n, kc, nout, ks) in [(96, 96, 96, 1, 2, 32, 3), (96, 96, 96, 4, 4, 32, 3), (48, 48, 48, 1, 2, 64, 3),
                                        (24, 24, 24, 2, 8, 128, 3), (12, 12, 12, 1, 16, 256, 3), (6, 6, 6, 1, 32, 512, 3),
                                        (6, 6, 6, 1, 32, 2048, 1), (96, 96, 96, 1, 2, 16, 1), (128, 128, 128, 2, 2, 32, 3)]:
        t = plan_conv(X, Y, Z, n, kc, nout, ks)
        a = _lib.ConvArgs()
        a.n_img, a.Z, a.Y, a.X = n, Z, Y, X
        a.src_cbt, a.ksize, a.n_kchunks = 2 * kc, ks, kc
        a.NT, a.n_ntiles, a.TX, a.TY, a.TZ, a.stages = t.NT, t.n_ntiles, t.TX, t.TY, t.TZ, t.stages
        a.out_mode, a.out_channels = 0, nout
        for i in range(kc):
            a.a_cb[i] = 2 * i
        got = _lib.lib.mmseg_conv3d_smem_bytes(C.byref(a))
        assert got == t.smem_bytes, (X, ks, nout, got, t, _lib.last_error())
        assert _lib.lib.mmseg_conv3d_tiles_per_img(C.byref(a)) == t.tiles_per_img


def test_rolling_z_plans_match_library():
    """plan_roll (rolling-z kernel: z segments, resident weights, optional paired K chunks) must be accepted by the C-side
    plan with the same shared-memory size, and its items must cover the volume."""
    import ctypes as C
    from mmseg_b200 import _lib
    from mmseg_b200.tiling import plan_roll, ROLL_FLAG, ROLL_KPAIR_FLAG
    for (X, Y, Z, n, kc, kpb) in [(96, 96, 96, 8, 1, 1), (96, 96, 96, 8, 2, 1), (96, 96, 96, 8, 2, 2), (96, 96, 96, 8, 4, 2),
                                  (128, 128, 128, 2, 4, 2), (128, 128, 128, 2, 2, 2), (48, 40, 36, 2, 1, 1), (20, 12, 30, 1, 2, 2)]:
        t = plan_roll(X, Y, Z, n, kc, 32, kpb)
        assert t is not None and t.roll and t.kpb == kpb and t.mt == 1 and t.NT == 32 and t.n_ntiles == 1
        assert (t.TY - 1) * (t.TX + 2) + t.TX <= 128
        assert t.tiles_per_img == -(-X // t.TX) * -(-Y // t.TY) * -(-Z // t.TZ)
        a = _lib.ConvArgs()
        a.n_img, a.Z, a.Y, a.X = n, Z, Y, X
        a.src_cbt, a.ksize, a.n_kchunks = 2 * kc, 3, kc
        a.NT, a.n_ntiles, a.TX, a.TY, a.TZ, a.stages = 32, 1, t.TX, t.TY, t.TZ, t.stages
        a.out_mode, a.out_channels = 0, 32
        a.flags = ROLL_FLAG | (ROLL_KPAIR_FLAG if kpb == 2 else 0)
        for i in range(kc):
            a.a_cb[i] = 2 * i
        got = _lib.lib.mmseg_conv3d_smem_bytes(C.byref(a))
        assert got == t.smem_bytes, (X, kc, kpb, got, t, _lib.last_error())
        assert _lib.lib.mmseg_conv3d_tiles_per_img(C.byref(a)) == t.tiles_per_img
    # not eligible: C_out != 32; rejected by the library: paired chunks that are not adjacent channel blocks
    assert plan_roll(96, 96, 96, 8, 2, 64) is None
    a.n_kchunks, a.flags = 2, ROLL_FLAG | ROLL_KPAIR_FLAG
    a.a_cb[0], a.a_cb[1] = 0, 4
    a.src_cbt = 8
    assert _lib.lib.mmseg_conv3d_smem_bytes(C.byref(a)) < 0 and "adjacent" in _lib.last_error()


def test_norm_option_tables_host_logic():
    """engine.norm_kind / ConvRunner._static_table: the (mean, rstd) + shift tables handed to the apply kernel reproduce
    BatchNorm3d in eval mode (running statistics + affine) and Identity; train-mode BatchNorm is refused loudly."""
    import pytest
    import torch
    import torch.nn.functional as F
    from mmseg_b200.engine import ConvRunner, norm_kind
    assert norm_kind(None) == "instance" and norm_kind(torch.nn.InstanceNorm3d(8)) == "instance"
    assert norm_kind(torch.nn.GroupNorm(8, 16)) == "group" and norm_kind(torch.nn.Identity()) == "none"
    bn = torch.nn.BatchNorm3d(16)
    with pytest.raises(NotImplementedError):
        norm_kind(bn)                                   # training mode: batch statistics are not built
    with pytest.raises(NotImplementedError):
        norm_kind(torch.nn.LayerNorm(4))
    bn.eval()
    assert norm_kind(bn) == "batch"
    torch.manual_seed(0)
    with torch.no_grad():
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.normal_(1.0, 0.3)
        bn.bias.normal_(0.0, 0.3)
    r = ConvRunner(False, torch.device("cpu"))
    mr, shift = r._static_table(bn, "batch", 2, 16, torch.device("cpu"))
    x = torch.randn(2, 16, 3, 4, 5)
    got = (x - mr[:, :, 0].view(2, 16, 1, 1, 1)) * mr[:, :, 1].view(2, 16, 1, 1, 1) + shift.view(2, 16, 1, 1, 1)
    want = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=False, eps=bn.eps)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    mr2, shift2 = r._static_table(bn, "batch", 2, 16, torch.device("cpu"))
    assert mr2 is mr                                    # cached until a parameter / buffer changes
    with torch.no_grad():
        bn.bias.add_(1.0)
    mr3, shift3 = r._static_table(bn, "batch", 2, 16, torch.device("cpu"))
    assert torch.allclose(shift3, shift + 1.0)
    mrn, shn = r._static_table(torch.nn.Identity(), "none", 3, 8, torch.device("cpu"))
    assert shn is None and bool((mrn[:, :, 0] == 0).all()) and bool((mrn[:, :, 1] == 1).all())


def test_rolling_z_schedule_invariants():
    """The stage stream of the rolling-z kernel (restated in tests/emulate.py): every output plane receives exactly the
    taps dz of its (up to three) in-volume input planes from every K chunk, is acquired before its first MMA, is signalled
    complete exactly once and only after its last contribution, and a ring slot is never reused before it completed."""
    from tests.emulate import roll_schedule
    R = 16
    for (Z, ZS, kc_n, kpb) in [(96, 32, 2, 1), (96, 32, 4, 2), (96, 48, 1, 1), (40, 17, 2, 2), (37, 37, 2, 1), (1, 4, 1, 1),
                               (5, 2, 4, 2), (128, 32, 4, 2), (19, 7, 4, 2)]:
        st = roll_schedule(Z, ZS, kc_n, kpb, R, n_items_xy=3)
        n_seg = -(-Z // ZS)
        contrib, acquired, completed = {}, set(), {}
        for i, s in enumerate(st):
            assert 0 <= s["z"] < Z                                   # only in-volume input planes are loaded
            for g in s["acquire"]:
                assert g not in acquired
                acquired.add(g)
                assert g < R or (g - R) in completed, (g, "slot reused before its previous plane completed")
            zo_first = None
            for (col, n, dz_first) in s["mma"]:
                assert 1 <= n <= 3 and col + n <= R                  # one MMA never wraps the ring
                for j in range(n):
                    dz = dz_first - j                                # weight rows are dz-descending
                    zo = s["q"] - dz                                 # output plane inside the segment
                    assert 0 <= zo < s["zsv"] and 0 <= dz <= 2
                    g = s["gz"] + zo
                    assert g % R == col + j and g in acquired and g not in completed
                    for c in range(s["kc"], s["kc"] + s["chunks"]):
                        key = (g, dz, c)
                        assert key not in contrib
                        contrib[key] = i
                    zo_first = zo if zo_first is None else zo_first
            for g in s["complete"]:
                assert g not in completed
                completed[g] = i
        total_planes = 3 * Z                                         # three columns of the same depth
        assert sorted(completed) == list(range(total_planes)) and acquired == set(range(total_planes))
        for g in range(total_planes):
            z = g % Z                                                # global plane -> volume z (segments tile the depth)
            want = {(g, dz, c) for dz in range(3) for c in range(kc_n) if 0 <= z + dz - 1 < Z}
            got = {k for k in contrib if k[0] == g}
            assert got == want, (Z, ZS, g, sorted(want - got), sorted(got - want))
            assert max(contrib[k] for k in got) <= completed[g]       # complete only after the last contribution


def test_epilogue_incremental_row_mapping():
    """The conv epilogue advances its voxel position by 128 accumulator rows per M tile without dividing (conv_tc.cu:
    xx += dx128, carry into yy, yy += dy128, carry into zz for the flat k=1 tiles): the carries are exact for every plane
    shape the planner can produce."""
    for PX in range(1, 129):
        for PY in (1, 2, 3, 4, 5, 7, 8, 11, 16, 31, 64):
            pxy = PX * PY
            dz128 = 128 // pxy
            r = 128 - dz128 * pxy
            dy128, dx128 = r // PX, r % PX
            for L0 in (0, 1, 31, 32, 63, 95, 96, 127):
                zz = L0 // pxy
                yy, xx = (L0 - zz * pxy) // PX, (L0 - zz * pxy) % PX
                for m in range(8):
                    L = m * 128 + L0
                    assert (zz, yy, xx) == (L // pxy, (L % pxy) // PX, (L % pxy) % PX), (PX, PY, L)
                    xx += dx128
                    if xx >= PX:
                        xx -= PX
                        yy += 1
                    yy += dy128
                    if yy >= PY:
                        yy -= PY
                        zz += 1
                    zz += dz128


def test_k1_gemm_tiles_avoid_the_spilling_odd_m_tile_counts():
    """conv3d_tc_kernel<MT, 1> spills 340 B for MT = 3, 5, 7 (conv_tc.ptxas.log): the planner never picks them."""
    from mmseg_b200.tiling import plan_conv
    for (X, n, kc, nout) in [(48, 4, 3, 144), (48, 4, 12, 48), (24, 4, 6, 288), (24, 1, 24, 96), (96, 4, 6, 48), (12, 2, 12, 576),
                             (6, 4, 24, 1152), (3, 1, 48, 768), (48, 8, 4, 256)]:
        t = plan_conv(X, X, X, n, kc, nout, 1)
        assert t.mt not in (3, 5, 7), (X, n, kc, nout, t)


def test_wgrad_channel_groups_and_alignment_rule():
    import ctypes as C
    from mmseg_b200 import _lib
    from mmseg_b200 import kernels as K
    # UNet widths keep (32, 32); SwinUNETR's 48 * 2^s widths get 24- / 48-channel groups instead of 16
    assert K.wgrad_groups(3, [32], 32) == (32, 32) and K.wgrad_groups(3, [64, 64], 64) == (32, 32)
    assert K.wgrad_groups(3, [48], 48) == (24, 48) and K.wgrad_groups(3, [96, 96], 96) == (32, 48)
    assert K.wgrad_groups(1, [48], 144) == (48, 48) and K.wgrad_groups(1, [192], 48) == (96, 48)
    assert K.wgrad_groups(3, [16], 32) == (16, 32)

    def smem(TX, TY, cig, cot, ks=3, X=16, Y=16, Z=8):
        a = _lib.WgradArgs()
        a.n_img, a.Z, a.Y, a.X, a.ksize = 1, Z, Y, X, ks
        a.TX, a.TY, a.TZ = TX, TY, 8
        a.cig_blocks, a.cot_blocks, a.n_cig, a.n_cot = cig, cot, 1, 1
        a.x_cbt, a.y_cbt, a.y_cb0, a.n_part = cig, cot, 0, 1
        return _lib.lib.mmseg_conv3d_wgrad_smem_bytes(C.byref(a))

    assert smem(16, 8, 4, 4) > 0
    # 3 channel blocks x (8 + 2) x 10 voxels x 16 B = 4800 B per x-shifted copy: not a multiple of 128 -> rejected (TMA)
    assert smem(10, 8, 3, 6) < 0 and "128-byte" in _lib.last_error()
    assert smem(12, 12, 3, 6) > 0
    # whatever the planner returns for the widths in use is accepted by the library
    for (X, cig, cot, ks) in [(96, 3, 6, 3), (10, 3, 6, 3), (6, 3, 6, 3), (128, 4, 4, 3), (48, 6, 6, 1), (24, 12, 6, 1)]:
        TX, TY, TZ = K._plan_wgrad_tile(X, X, X, ks, cig, cot)
        assert smem(TX, TY, cig, cot, ks, X, X, X) > 0, (X, cig, cot, ks, TX, TY)


def test_wgrad_plan_rows_groups_and_split():
    """The cached host plan of a wgrad launch: accumulator-row map of concat / padded inputs, split-K depth."""
    from mmseg_b200 import kernels as K
    plan = lambda segs, S, n, cout, ks=3: K._wgrad_plan(tuple(segs), S, S, S, n, cout, ks, "auto", "0")
    # decoder concat [up 32 | skip 32] -> two 32-row groups at channel blocks 0 and 4, rows in concat order
    cin, cig, ntc, groups, ci_map, n_cig, n_cot, TX, TY, TZ, n_part = plan([(0, 32), (32, 32)], 32, 2, 32)
    assert (cin, cig, ntc, groups, n_cig, n_cot) == (64, 32, 32, (0, 4), 2, 1) and ci_map == tuple(range(64))
    assert (TZ, n_part) == (8, 32)        # 32 tiles of 8 planes: fewer tiles than the 74 persistent CTAs per group pair
    # first layer: 2 real channels padded to one 16-row group
    cin, cig, ntc, groups, ci_map, n_cig, n_cot, *_ = plan([(0, 2)], 16, 1, 32)
    assert (cin, cig, groups, ci_map, n_cig) == (2, 16, (0,), (0, 1), 1)
    # two 8-channel segments stored 16 apart (each padded to a 16-row group): rows 0-7 and 16-23
    cin, cig, ntc, groups, ci_map, *_ = plan([(0, 8), (16, 8)], 8, 1, 16)
    assert (cin, cig, groups) == (16, 16, (0, 2)) and ci_map == tuple(range(8)) + tuple(range(16, 24))
    # SwinUNETR width 48: 24-row groups; a deep layer (768 -> 768 at 3^3) is one partial per channel-group pair
    cin, cig, ntc, groups, ci_map, n_cig, n_cot, TX, TY, TZ, n_part = plan([(0, 768)], 3, 2, 768)
    assert (cig, ntc, n_cig, n_cot, n_part) == (32, 48, 24, 16, 1)
    assert plan([(0, 48)], 48, 1, 48)[1:4] == (24, 48, (0, 3))
    assert plan([(0, 64)], 8, 1, 256, ks=1)[1:3] == (64, 256)


def test_capture_guard_disables_gc_and_restores_it():
    import gc
    from mmseg_b200.src.trainer.inference import capture_guard
    assert gc.isenabled()
    with capture_guard():
        assert not gc.isenabled()
    assert gc.isenabled()
    gc.disable()
    try:
        with capture_guard():
            assert not gc.isenabled()
        assert not gc.isenabled()       # it was off before: stays off
    finally:
        gc.enable()
