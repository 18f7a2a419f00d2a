"""GPU parity tests of the SwinUNETR path (marker `gpu`): swin.cu kernels and the whole model through the C ABI vs the
CPU oracle (PARITY UNPINNED against MONAI, see oracle/swin_unetr.py) / torch fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tests.gpu_cases_swin  # noqa: F401
    from mmseg_b200 import _lib
    _lib.require_device()


def _c():
    import tests.gpu_cases_swin as c
    return c


def test_patch_embed():
    _c().patch_embed_case()
    _c().patch_embed_case(cin=4, feat=96, dims=(3, 5, 4), n=1)


@pytest.mark.parametrize("kw", [
    dict(), dict(with_add=False, affine=False), dict(channels=384, dims=(3, 3, 3), n=1), dict(mode="bf16"),
    dict(channels=48, c0=48, with_add=False, affine=False),
])
def test_layernorm_residual(kw):
    _c().layernorm_case(**kw)


def test_patch_merging_layernorm():
    _c().merge_case()
    _c().merge_case(channels=96, dims=(6, 6, 6), n=1, mode="bf16")


@pytest.mark.parametrize("kw", [
    dict(dims=(8, 9, 10), heads=3, shift=False),        # padding to 14^3, no shift
    dict(dims=(8, 9, 10), heads=3, shift=True),         # padding + shift mask
    dict(dims=(14, 7, 16), heads=2, shift=True, n=1),   # one axis equal to the window: that axis is not shifted
    dict(dims=(6, 6, 6), heads=6, shift=True, n=1),     # window shrinks to 6^3 (216 tokens), shift cancelled, index[:n, :n]
    dict(dims=(3, 3, 3), heads=12, shift=False, n=2),   # 27 tokens
    dict(dims=(16, 16, 16), heads=3, shift=True, n=1, mode="bf16"),
])
def test_window_attention(kw):
    _c().window_attention_case(**kw)


def test_residual_instnorm_act():
    _c().resnorm_case()
    _c().resnorm_case(identity=True)
    _c().resnorm_case(channels=96, dims=(3, 3, 3), n=1, mode="bf16")


@pytest.mark.parametrize("size,mode", [(64, "fp16"), (64, "bf16"), (96, "fp16"), ((64, 96, 128), "fp16")])
def test_swin_unetr_vs_oracle(size, mode):
    _c().swin_unetr_case(size=size, mode=mode)


def test_swin_unetr_sliding_window():
    _c().swin_sliding_window_case()


@pytest.mark.parametrize("kw", [dict(), dict(with_add=False), dict(channels=384, dims=(3, 3, 3), n=1)])
def test_layernorm_backward(kw):
    _c().ln_backward_case(**kw)


def test_gelu_merge_patch_embed_backward():
    _c().gelu_merge_patch_backward_case()


@pytest.mark.parametrize("kw", [dict(), dict(shift=False), dict(dims=(6, 6, 6), heads=3, shift=True),
                                dict(dims=(14, 14, 14), heads=1, shift=True, n=2)])
def test_window_attention_backward(kw):
    _c().window_attention_backward_case(**kw)


@pytest.mark.parametrize("cin,cout", [(32, 48), (48, 48)])
def test_unet_res_block_backward(cin, cout):
    _c().res_block_backward_case(cin, cout)


@pytest.mark.parametrize("size,n", [(64, 1), ((64, 32, 96), 2)])
def test_swin_unetr_training_step_vs_fp64_autograd(size, n):
    _c().swin_train_step_case(size=size, n=n)


def test_swin_unetr_trainer_train_validate_predict(tmp_path):
    _c().swin_trainer_case(tmp_path)
