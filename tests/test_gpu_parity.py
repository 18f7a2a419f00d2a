"""GPU parity tests (marker `gpu`): every case calls the CUDA kernels through the C ABI (ctypes) and compares with the
CPU oracle / torch fp64 on the same seeded inputs, or with the golden vectors generated from the reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tests.gpu_cases  # noqa: F401  (loads libmmseg_b200.so)
    from mmseg_b200 import _lib
    _lib.require_device()


def _c():
    import tests.gpu_cases as c
    return c


@pytest.mark.parametrize("args,kw", [
    ((16, 32, (8, 12, 20)), {}),
    ((2, 32, (6, 10, 24)), {}),
    ((32, 32, (7, 9, 30)), dict(n_img=2, tile=(30, 4, 2))),      # ragged tiles in x, y, z
    ((32, 64, (5, 11, 50)), dict(tile=(25, 5, 2))),
    ((64, 128, (6, 6, 6)), dict(n_img=2)),
    ((128, 256, (6, 6, 6)), {}),
    ((32, 32, (16, 24, 96)), {}),
    ((32, 32, (6, 8, 20)), dict(split=True)),
    ((64, 32, (8, 16, 48)), dict(split=True)),
    ((32, 16, (5, 6, 20)), dict(ks=1)),
    ((256, 512, (3, 3, 3)), {}),
])
def test_conv3d_tcgen05(args, kw):
    _c().conv_case(*args, **kw)


@pytest.mark.parametrize("args,kw", [
    ((16, 32, (8, 12, 16)), {}),
    ((2, 32, (8, 12, 16)), dict(split=True)),
    ((32, 64, (8, 8, 16)), dict(n_img=2, slope=0.2, split=True)),
])
def test_conv_norm_act_pool_block(args, kw):
    _c().conv_block_case(*args, **kw)


@pytest.mark.parametrize("args,kw", [((64, (3, 5, 6)), {}), ((32, (4, 4, 12)), dict(n_img=2, split=True))])
def test_conv_transpose_k2s2(args, kw):
    _c().convt_case(*args, **kw)


@pytest.mark.parametrize("split", [False, True])
def test_logits_1x1(split):
    _c().logits_case(32, 8, (6, 7, 20), split=split)


def test_pack_unpack_roundtrip():
    _c().pack_roundtrip_case()


@pytest.mark.parametrize("features,S,n,mode", [((16, 32, 64), 32, 1, "parity"), ((16, 32, 64), 32, 2, "bf16"),
                                                ((32, 64, 128, 256, 512), 96, 1, "parity"),
                                                ((32, 64, 128, 256, 512), 96, 1, "bf16")])
def test_unet3d_vs_oracle(features, S, n, mode):
    _c().unet_case(features, S, n, mode)


@pytest.mark.parametrize("mode", ["parity", "bf16"])
def test_unet3d_golden_from_reference(mode):
    _c().unet_golden_case(mode)


@pytest.mark.parametrize("kw", [dict(), dict(weights=True), dict(include_background=False), dict(B=1, C=3, shape=(5, 7, 9)),
                                dict(B=1, C=8, shape=(64, 64, 64))])
def test_dicece(kw):
    _c().dicece_case(**kw)


def test_dicece_golden():
    _c().dicece_golden_case()


def test_swi_blend_finalize_exact():
    _c().swi_constant_predictor_case()


@pytest.mark.parametrize("kw", [dict(mode="gaussian"), dict(mode="constant"), dict(vol_shape=(33, 64, 40), overlap=0.25),
                                dict(nmode="bf16")])
def test_sliding_window_vs_oracle(kw):
    _c().swi_case(**kw)
