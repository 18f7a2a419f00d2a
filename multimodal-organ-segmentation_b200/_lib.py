"""ctypes binding of libmmseg_b200.so (the C ABI declared in include/mmseg_b200.h).

There is no fallback: if the shared library is missing the import of this module raises, and every compute entry
point raises RuntimeError when the call is rejected or no sm_100 device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmseg_b200.so")

MAX_KCHUNKS = 256

OUT_BLOCKED_BF16 = 0
OUT_BLOCKED_F32 = 1
OUT_BLOCKED_BF16_HILO = 2
OUT_CONVT_K2S2 = 3
OUT_NCDHW_F32 = 4

FMT_BF16 = 0
FMT_FP16 = 1
CONV_FP16_FLAG = 64   # MMSEG_CONV_FP16


class ConvArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("weights", C.c_void_p), ("bias", C.c_void_p), ("dst", C.c_void_p),
        ("stats_partial", C.c_void_p),
        ("n_img", C.c_int32), ("Z", C.c_int32), ("Y", C.c_int32), ("X", C.c_int32),
        ("src_cbt", C.c_int32), ("ksize", C.c_int32), ("n_kchunks", C.c_int32),
        ("NT", C.c_int32), ("n_ntiles", C.c_int32),
        ("TX", C.c_int32), ("TY", C.c_int32), ("TZ", C.c_int32),
        ("stages", C.c_int32), ("out_mode", C.c_int32), ("out_channels", C.c_int32),
        ("dst_cbt", C.c_int32), ("dst_cb_off", C.c_int32), ("dst_lo_off", C.c_int32),
        ("flags", C.c_int32),
        ("a_cb", C.c_int16 * MAX_KCHUNKS),
    ]


MAX_WGRAD_GROUPS = 64


class WgradArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("partial", C.c_void_p),
        ("n_img", C.c_int32), ("Z", C.c_int32), ("Y", C.c_int32), ("X", C.c_int32),
        ("ksize", C.c_int32), ("TX", C.c_int32), ("TY", C.c_int32), ("TZ", C.c_int32),
        ("cig_blocks", C.c_int32), ("cot_blocks", C.c_int32), ("n_cig", C.c_int32), ("n_cot", C.c_int32),
        ("x_cbt", C.c_int32), ("y_cbt", C.c_int32), ("y_cb0", C.c_int32), ("n_part", C.c_int32),
        ("x_cb", C.c_int16 * MAX_WGRAD_GROUPS),
    ]


class AdamwTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("step", C.c_void_p),
                ("n", C.c_int64)]


class RepackDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("n_off", C.c_void_p), ("k_off", C.c_void_p), ("dst", C.c_void_p),
                ("n_out", C.c_int32), ("NT", C.c_int32), ("n_kc", C.c_int32), ("n_kc_total", C.c_int32),
                ("ksize", C.c_int32), ("flip", C.c_int32), ("hi_copies", C.c_int32), ("has_lo", C.c_int32),
                ("first_block", C.c_int64)]


ADAMW_CHUNK = 8192


class NormBwdArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("mean_rstd", C.c_void_p), ("gA", C.c_void_p), ("gP", C.c_void_p),
        ("partial", C.c_void_p), ("dx", C.c_void_p), ("chan_scale", C.c_void_p), ("chan_bias", C.c_void_p),
        ("n_img", C.c_int32), ("cb", C.c_int32), ("Z", C.c_int32), ("Y", C.c_int32), ("X", C.c_int32),
        ("gA_cbt", C.c_int32), ("gA_cb_off", C.c_int32), ("gP_cbt", C.c_int32), ("gP_cb_off", C.c_int32),
        ("dx_cbt", C.c_int32), ("dx_cb_off", C.c_int32), ("n_chunks", C.c_int32),
        ("gA_scale", C.c_float), ("slope", C.c_float), ("m12", C.c_void_p),
    ]


class NormArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("mean_rstd", C.c_void_p), ("dst", C.c_void_p), ("pooled", C.c_void_p),
        ("n_img", C.c_int32), ("cb", C.c_int32), ("Z", C.c_int32), ("Y", C.c_int32), ("X", C.c_int32),
        ("src_is_f32", C.c_int32),
        ("dst_cbt", C.c_int32), ("dst_cb_off", C.c_int32), ("dst_lo_off", C.c_int32),
        ("pool_cbt", C.c_int32), ("pool_cb_off", C.c_int32), ("pool_lo_off", C.c_int32),
        ("slope", C.c_float),
        ("stats_partial", C.c_void_p), ("mean_rstd_out", C.c_void_p), ("tiles_per_img", C.c_int32), ("eps", C.c_float),
        ("shift", C.c_void_p), ("act", C.c_int32), ("elem_fmt", C.c_int32),
    ]


class SwinAttnArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("out", C.c_void_p), ("table", C.c_void_p), ("qkv_bias", C.c_void_p),
        ("n_img", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("window", C.c_int32 * 3), ("shift", C.c_int32 * 3),
        ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("qkv_cbt", C.c_int32), ("out_cbt", C.c_int32), ("out_cb_off", C.c_int32),
        ("scale", C.c_float), ("elem_fmt", C.c_int32), ("lse", C.c_void_p),
    ]


# every symbol include/mmseg_b200.h declares: name -> (restype, argtypes)
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = {
    "mmseg_version": (C.c_int, []),
    "mmseg_last_error": (C.c_char_p, []),
    "mmseg_device_ok": (C.c_int, []),
    "mmseg_sizeof": (C.c_int, [C.c_int]),
    "mmseg_conv3d_fwd": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "mmseg_conv3d_smem_bytes": (_i64, [C.POINTER(ConvArgs)]),
    "mmseg_conv3d_tiles_per_img": (_i32, [C.POINTER(ConvArgs)]),
    "mmseg_conv3d_wgrad": (C.c_int, [C.POINTER(WgradArgs), _vp]),
    "mmseg_conv3d_wgrad_smem_bytes": (_i64, [C.POINTER(WgradArgs)]),
    "mmseg_wgrad_reduce": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "mmseg_instnorm_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _f32, _vp, _vp]),
    "mmseg_instnorm_act_apply": (C.c_int, [C.POINTER(NormArgs), _vp]),
    "mmseg_instnorm_act_bwd_reduce": (C.c_int, [C.POINTER(NormBwdArgs), _vp]),
    "mmseg_instnorm_act_bwd_apply": (C.c_int, [C.POINTER(NormBwdArgs), _vp]),
    "mmseg_modality_dot": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _i32, _vp, _vp]),
    "mmseg_unshuffle_k2s2": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mmseg_im2col_k3_c1": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mmseg_pack_ncdhw": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_unpack_ncdhw": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swi_gather": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swi_gather_ncdhw": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mmseg_swi_blend": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _f32, _vp, _vp,
                                  _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swi_finalize": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i32, _vp, _vp]),
    "mmseg_swi_add_partial": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i64, _vp]),
    "mmseg_dicece_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _f32, _f32, _f32, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "mmseg_dicece_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _f32, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mmseg_tversky_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _f32, _f32, _f32, _vp, _i32, _vp, _vp, _vp]),
    "mmseg_tversky_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "mmseg_focal": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, _f32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mmseg_cross_attention_fwd": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i64,
                                            _f32, _vp, _vp]),
    "mmseg_cross_attention_bwd": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp,
                                            _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _f32, _vp]),
    "mmseg_add_stats": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _i32, _vp]),
    "mmseg_confusion_hist": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "mmseg_channel_mean": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _i32, _vp, _i32, _vp]),
    "mmseg_gate_mlp": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mmseg_gate_mlp_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmseg_modality_combine": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _f32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_modality_max": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i64, _vp, _i32, _i32, _vp]),
    "mmseg_swi_logits_blend": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp,
                                         _vp, _f32, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_pack_ncdhw_ex": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _i32, _vp]),
    "mmseg_groupnorm_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i64, _f32, _vp, _vp, _vp, _vp, _vp]),
    "mmseg_trilinear_resize": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp]),
    "mmseg_conv1x1_logits": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _i32, _vp, _i32, _vp]),
    "mmseg_channel_stats": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _vp, _vp]),
    "mmseg_modality_normalize": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mmseg_weights_repack": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "mmseg_weights_repack_multi": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _f32, _vp]),
    "mmseg_gather_f32": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "mmseg_adamw_multi": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "mmseg_swin_patch_embed": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swin_layernorm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _i32, _f32, _i32, _vp]),
    "mmseg_swin_merge_ln": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "mmseg_swin_window_attention": (C.c_int, [C.POINTER(SwinAttnArgs), _vp]),
    "mmseg_instnorm_residual_act": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i64,
                                              _f32, _i32, _vp]),
    "mmseg_swin_ln_fwd_train": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _f32, _vp]),
    "mmseg_swin_ln_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp]),
    "mmseg_swin_ln_param_grad": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp]),
    "mmseg_gelu_bwd": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "mmseg_lrelu_mask_mul": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _vp]),
    "mmseg_swin_merge_gather": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swin_merge_scatter": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swin_patch_embed_wgrad": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mmseg_swin_window_attention_bwd": (C.c_int, [C.POINTER(SwinAttnArgs), _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mmseg_maxpool3d_2": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C multimodal-organ-segmentation_b200/csrc`). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)  # AttributeError if the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


# the ctypes structures must have the layout the library was compiled with (a field appended on one side only would shift
# every later argument silently)
for _which, _struct in enumerate((ConvArgs, WgradArgs, NormArgs, NormBwdArgs, AdamwTensor, RepackDesc, SwinAttnArgs)):
    if lib.mmseg_sizeof(_which) != C.sizeof(_struct):
        raise ImportError(f"{_struct.__name__}: ctypes layout ({C.sizeof(_struct)} bytes) differs from libmmseg_b200.so "
                          f"({lib.mmseg_sizeof(_which)} bytes): rebuild the library (make -C multimodal-organ-segmentation_b200/csrc)")


def last_error() -> str:
    return lib.mmseg_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


_device_ok = None


def require_device() -> None:
    """Fail loudly when there is no B200-class device: the product path has no CPU route."""
    global _device_ok
    if _device_ok is None:
        _device_ok = bool(lib.mmseg_device_ok())
    if not _device_ok:
        raise RuntimeError("mmseg_b200 needs a CUDA device of compute capability 10.x (sm_100a); none is visible "
                           "and there is no CPU fallback")
