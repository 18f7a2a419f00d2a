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
    ((32, 32, (6, 8, 20)), dict(split="fp16")),                   # fp16 operands, fp16 raw output
    ((32, 64, (5, 11, 50)), dict(split="fp16", raw_f32=True, tile=(25, 5, 2))),
    ((64, 32, (8, 16, 48)), dict(split="fp16a2")),                # [A_hi | A_lo] x [W | W]
    ((32, 32, (6, 8, 20)), dict(split="fp16w2")),                 # [A | A] x [W_hi | W_lo]
    ((2, 32, (6, 10, 24)), dict(split="fp16x3")),
    ((32, 16, (5, 6, 20)), dict(ks=1, split="fp16x3")),
])
def test_conv3d_tcgen05(args, kw):
    _c().conv_case(*args, **kw)


@pytest.mark.parametrize("args,kw", [
    ((16, 32, (20, 12, 30)), dict(roll="auto")),
    ((2, 32, (9, 11, 27)), dict(roll="auto", n_img=3)),
    ((32, 32, (40, 7, 26)), dict(n_img=2, roll=(13, 4, 17, 8))),     # 3 segments (17, 17, 6), ring wraps, 2 K chunks
    ((64, 32, (19, 6, 20)), dict(roll=(20, 5, 7, 12))),              # 4 K chunks, segments 7, 7, 5
    ((16, 32, (1, 5, 8)), dict(roll=(8, 5, 4, 8))),                  # a single output plane
    ((32, 32, (37, 5, 24)), dict(roll=(24, 5, 37, 10), n_img=2)),    # one long segment: 37 planes through a 16-slot ring
    ((32, 32, (40, 7, 26)), dict(n_img=2, roll=(13, 4, 17, 6, 2))),  # two K chunks per TMA stage (4-block box)
    ((64, 32, (19, 6, 20)), dict(roll=(20, 5, 7, 5, 2))),            # 4 K chunks = 2 paired stages per plane, 5-stage ring
    ((32, 32, (40, 7, 26)), dict(n_img=2, roll=(13, 4, 17, 8), raw_f32=True)),              # fp32 raw output (bf16 operands)
    ((64, 32, (19, 6, 20)), dict(roll=(20, 5, 7, 5, 2), split="fp16", raw_f32=True)),        # fp16 operands, fp32 raw output
    ((32, 32, (20, 12, 30)), dict(roll="auto", split="fp16")),                               # fp16 operands and output
])
def test_conv3d_rolling_z(args, kw):
    """conv3d_roll_kernel (TMEM ring of output planes) vs F.conv3d, raw output and InstanceNorm partial sums."""
    _c().conv_case(*args, **kw)


@pytest.mark.parametrize("mode,in_ch", [("fp16m", 2), ("parity", 2), ("fp16i", 4), ("fp16x3", 1)])
def test_first_layer_packed_split(mode, in_ch):
    _c().in_packed_case(mode=mode, in_ch=in_ch)


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


@pytest.mark.parametrize("args,kw", [
    ((32, 8, (6, 7, 20)), {}),
    ((32, 8, (5, 3, 7)), dict(split=True)),              # 105 voxels: ragged 4-voxel groups
    ((16, 3, (4, 5, 9)), dict(n_img=1)),                 # COUT template 4
    ((64, 14, (3, 4, 10)), dict(split=True, c0=16)),     # COUT template 16, channel offset inside a wider buffer
])
def test_logits_1x1_cuda_cores(args, kw):
    _c().logits_cuda_core_case(*args, **kw)


@pytest.mark.parametrize("a,b", [((3, 4, 5), (6, 8, 10)), ((6, 5, 7), (12, 10, 14)), ((4, 4, 4), (4, 9, 1)),
                                 ((1, 3, 300), (5, 2, 333))])
def test_trilinear_resize_align_corners(a, b):
    _c().trilinear_case(a, b)


def test_deep_supervision_head_golden():
    _c().deep_supervision_golden_case()


@pytest.mark.parametrize("norm,mode", [("group", "parity"), ("batch", "parity"), ("none", "parity"), ("group", "bf16"),
                                       ("batch", "bf16")])
def test_unet_norm_options_golden(norm, mode):
    _c().unet_norm_golden_case(norm, mode)


@pytest.mark.parametrize("kind", ["unet", "dual"])
def test_input_gradient(kind):
    _c().input_grad_case(kind)


def test_trainer_train_validate_resume_predict(tmp_path):
    _c().trainer_end_to_end_case(tmp_path)


@pytest.mark.parametrize("mode", ["parity", "bf16"])
def test_convblock_gelu_group_golden(mode):
    _c().convblock_gelu_group_golden_case(mode)


def test_suv_guided_attention_golden():
    _c().suv_guided_attention_golden_case()


def test_full_size_volume_properties():
    _c().full_volume_properties_case()


def test_full_size_train_step_properties():
    _c().full_train_step_properties_case()


def test_sliding_window_fused_head_bit_identical():
    _c().fused_head_case()


def test_predict_volume_host_to_host():
    _c().predict_volume_case()


def test_sliding_window_graphs_follow_weight_updates():
    _c().swi_weight_update_case()


@pytest.mark.parametrize("mode", ["parity", "fp16m", "bf16"])
def test_unet_odd_sizes_trilinear_resize_branch(mode):
    _c().unet_odd_size_case(mode=mode)


@pytest.mark.parametrize("norm", ["group", "batch", "none"])
def test_unet_training_with_non_instance_norms(norm):
    _c().unet_norm_training_case(norm)


def test_pack_unpack_roundtrip():
    _c().pack_roundtrip_case()


@pytest.mark.parametrize("features,S,n,mode", [((16, 32, 64), 32, 1, "parity"), ((16, 32, 64), 32, 2, "bf16"),
                                                ((16, 32, 64), 32, 2, "fp16"), ((16, 32, 64), 32, 1, "fp16w2"),
                                                ((16, 32, 64), 32, 1, "fp16a2"), ((16, 32, 64), 32, 2, "fp16x3"),
                                                ((16, 32, 64), 32, 2, "fp16m"), ((16, 32, 64), 32, 2, "fp16i"),
                                                ((32, 64, 128, 256, 512), 96, 1, "parity"),
                                                ((32, 64, 128, 256, 512), 96, 1, "fp16m"),
                                                ((32, 64, 128, 256, 512), 96, 1, "fp16"),
                                                ((32, 64, 128, 256, 512), 96, 1, "bf16")])
def test_unet3d_vs_oracle(features, S, n, mode):
    """Every rung of the numeric-mode ladder (numerics.py): the >= 16-bit modes against the north_star gates, the
    faster rungs against the noise class of their operand format (gpu_cases.NOISE_CLASS)."""
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
                                dict(nmode="bf16"), dict(nmode="fp16"), dict(nmode="fp16m"), dict(net="dual"),
                                dict(net="dual", nmode="fp16"), dict(net="dual", nmode="fp16m"),
                                dict(vol_shape=(50, 41, 70), mode="gaussian")])
def test_sliding_window_vs_oracle(kw):
    _c().swi_case(**kw)


# ------------------------------------------------------------------------------------------------ backward / training
@pytest.mark.parametrize("args,kw", [
    ((32, 32, (4, 8, 16)), {}),
    ((32, 32, (9, 12, 20)), dict(n_img=2)),                       # ragged tiles, several images
    ((16, 32, (8, 8, 8)), {}),
    ((2, 32, (6, 10, 24)), {}),                                   # first layer (2 real input channels)
    ((1, 32, (5, 9, 41)), dict(n_img=2)),                         # DualEncoder first layer, ragged x / y tiles
    ((2, 64, (3, 6, 70)), {}),                                    # 64 output channels: two passes of 8 warps
    ((64, 64, (8, 8, 8)), {}),
    ((64, 32, (6, 16, 16)), dict(segs=[(0, 32), (32, 32)])),      # concat input [up | skip]
    ((32, 8, (6, 7, 20)), dict(ks=1)),                            # out_conv
    ((64, 256, (4, 4, 8)), dict(ks=1)),                           # ConvTranspose GEMM view
    ((32, 32, (24, 24, 24)), dict(n_img=2)),
    ((256, 128, (4, 4, 4)), dict(n_img=2)),
    ((128, 128, (8, 8, 12)), {}),                                 # shallow split-K: tiled reduce (all 27 taps per block)
    ((96, 48, (6, 6, 6)), dict(n_img=2)),                         # SwinUNETR widths: 24-row / 48-column groups
    ((1, 16, (4, 6, 10)), dict(segs=[(8, 1)])),                   # one channel in block 1: im2col + k=1 wgrad path
])
def test_wgrad_tcgen05(args, kw):
    _c().wgrad_case(*args, **kw)


@pytest.mark.parametrize("kw", [dict(pool=False), dict(pool=True), dict(channels=32, shape=(6, 6, 8), n_img=1, slope=0.2, scale=0.5)])
def test_instnorm_act_pool_backward(kw):
    _c().norm_bwd_case(**kw)


@pytest.mark.parametrize("args", [(32, 32, (6, 8, 12), 1, 3), (64, 32, (5, 7, 9), 2, 3), (16, 32, (4, 6, 8), 1, 1)])
def test_dgrad_through_forward_kernel(args):
    _c().dgrad_case(*args)


def test_conv_transpose_backward():
    _c().convt_bwd_case()


@pytest.mark.parametrize("args,kw", [
    (("unet", (16, 32), 16, 1), {}),
    (("unet", (16, 32, 64), 16, 2), {}),
    (("dual", (16, 32), 16, 2), dict(fusion="late")),
    (("dual", (16, 32), 16, 1), dict(fusion="concat")),
    (("dual", (16, 32), 16, 1), dict(fusion="add", M=3)),
    (("dual", (16, 32), 16, 2), dict(fusion="attention")),
    (("dual", (16, 32), 16, 1), dict(fusion="attention", M=4)),
    # model.backbone.norm != instance through DualEncoder's fusions, and Dropout3d (the mask is reproduced from the seed)
    (("dual", (16, 32), 16, 2), dict(fusion="late", norm="group")),
    (("dual", (16, 32), 16, 2), dict(fusion="attention", norm="batch")),
    (("dual", (16, 32), 16, 2), dict(fusion="concat", norm="none")),
    (("dual", (16, 32), 16, 2), dict(fusion="add", norm="group", dropout=0.4)),
    (("unet", (16, 32), 16, 2), dict(norm="batch", dropout=0.4)),
    (("unet", (16, 32), 16, 2), dict(dropout=0.4)),
])
def test_training_step_vs_fp64_autograd(args, kw):
    _c().train_step_case(*args, **kw)


@pytest.mark.parametrize("name,fusion", [("dual_attention_2", "attention"), ("dual_concat_2", "concat"),
                                         ("dual_cross_attention_2", "cross_attention"), ("dual_add_2", "add"),
                                         ("dual_attention_4", "attention")])
def test_dual_encoder_golden_from_reference(name, fusion):
    _c().dual_golden_case(name)


# ------------------------------------------------------------------------------------------------ fusion modules / attention
def test_attention_modules_golden_from_reference():
    _c().cross_attention_golden_case()


@pytest.mark.parametrize("kw", [dict(C=64, shape=(8, 8, 8)), dict(C=128, shape=(12, 12, 12), n_img=1),
                                dict(C=256, shape=(6, 7, 9), n_img=1), dict(C=512, shape=(6, 6, 6), n_img=1),
                                dict(C=64, heads=2, shape=(5, 5, 5))])
def test_fused_cross_attention_vs_oracle(kw):
    _c().cross_attention_case(**kw)


def test_late_early_fusion_and_head():
    _c().late_early_head_case()


def test_focal_tversky_golden():
    _c().focal_tversky_golden_case()


# ------------------------------------------------------------------------------------------------ N1: trainer step glue
def test_weights_repack_matches_aten_packing():
    _c().weights_repack_case()


def test_fused_adamw_matches_torch():
    _c().fused_adamw_case()


def test_reference_cli_inference_on_dropin(tmp_path):
    """The reference's own main.py --mode inference, end to end on the drop-in packages (tests/dropin.py)."""
    _c().reference_cli_inference_case(tmp_path)


def test_device_transforms_match_reference_numpy():
    _c().device_transforms_case()


@pytest.mark.parametrize("kw", [dict(hd=32, n_tok=300), dict(hd=16, n_tok=128, heads=2), dict(hd=64, n_tok=513, n_img=1),
                                dict(hd=128, n_tok=260, heads=2, n_img=1), dict(hd=32, n_tok=1000, heads=1)])
def test_cross_attention_core_backward(kw):
    """dQ / dK / dV of the fused attention kernel vs fp64 autograd (ragged token counts, every head dim)."""
    _c().attention_core_bwd_case(**kw)


@pytest.mark.parametrize("kw", [dict(), dict(C=32, heads=4, shape=(4, 6, 5), n_img=1), dict(C=128, heads=4, shape=(8, 8, 8))])
def test_cross_attention_fusion_trains_through_the_kernels(kw):
    """Module-level gradient parity (inputs + all projections) vs fp64 autograd of the oracle maths."""
    _c().cross_attention_module_grad_case(**kw)


def test_bidirectional_cross_attention_trains_through_the_kernels():
    _c().bidirectional_attention_grad_case()
