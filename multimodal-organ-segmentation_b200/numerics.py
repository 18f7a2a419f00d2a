"""Numeric modes of the inference kernels — the "parity ladder".

Every contraction (Conv3d / ConvTranspose3d / 1x1) runs on tcgen05 `kind::f16` MMAs with fp32 accumulation in TMEM;
a mode chooses the 16-bit operand format and how many MMA passes approximate one fp32 product:

  name     operands  passes  products                                  operand significand (A, W)
  bf16     bf16      1       A*W                                       8, 8
  fp16     fp16      1       A*W                                       11, 11
  fp16w2   fp16      2       A*W_hi + A*W_lo                           11, 22   (weights exact, no extra activation traffic)
  fp16a2   fp16      2       A_hi*W + A_lo*W                           22, 11   (activations exact)
  parity   bf16      3       A_hi*W_hi + A_lo*W_hi + A_hi*W_lo         16, 16   (alias "bf16x3")
  fp16x3   fp16      3       same split in fp16                        22, 22
  fp16m    fp16      1 - 3   fp16 + more bits only where the error comes from (mixed mode, below): THE DEFAULT
  fp16i    fp16      1 - 2   the leanest mixed rung: input hi + lo, first-block weights split, fp32 raw outputs

Where the error of fp16 comes from (B200, 36-window bench crop vs the fp32 reference; tools/ladder_probe.py):
  * all of fp16:                                              3.2e-3 rel-L2, 99.73 % labels   (gate: 1e-3, 99.9 %)
  * exact activations everywhere (fp16a2): the weights alone  1.07e-3 — and half of THAT is the first ConvBlock3D
    (init_conv: only 54 / 864 products per output): its weights split -> 5.1e-4; every deeper level split -> 1.02e-3;
  * no activation split, first-block weights split, fp32 raw  2.7e-3;  + the INPUT volume stored hi + lo -> 9.1e-4:
    the fp16 rounding of the CT / PET intensities themselves is the dominant term;  + mid0 (the first conv's output)
    -> 7.4e-4;  + pool1 -> 6.8e-4;  every activation buffer -> 5.1e-4;
  * raw conv outputs in fp16 instead of fp32 on the full-resolution level alone: 6.8e-4 -> 1.4e-3.
`fp16m` = fp16 operands, fp32 raw conv outputs, the buffers {in, mid0, pool1} stored hi + lo and the weights of the
first block split: 6.8e-4 / 99.94 % / Dice 0.99938 at 1.7x the speed of the 3-pass split everywhere ("parity").

When the input is stored hi + lo AND the first conv's weights are split, the three passes of that thin layer (C_in = 2) are
packed into ONE 16-channel K chunk: virtual input channels [hi | lo | hi] against [W_hi | W_hi | W_lo] (engine.in_packed,
kernels.Blocked.packed_split) — the first layer then costs what it costs in single-pass fp16 (0.52 -> 0.22 ms per batch).

Split operands are extra K chunks of the same GEMM (kernels.a_chunk_table / pack_conv_weight), so every mode runs the
same kernels; `raw_f32` keeps the raw conv output (the InstanceNorm input) in fp32 instead of the 16-bit format.
The north_star gates (logits 2e-2 max-abs / 1e-3 rel-L2, labels >= 99.9 %, Dice within 1e-3) are measured per mode by
bench.py on the 36-window crop; training runs in bf16 (north_star: bf16 training step).
"""
import os
from dataclasses import dataclass, replace
from functools import lru_cache
from typing import Optional, Union

import torch

from . import _lib


def _match(tags, tag: str) -> bool:
    """tag "enc0.1" is selected by "enc0.1" or by its block "enc0"."""
    return any(tag == t or tag.startswith(t + ".") for t in tags)


@dataclass(frozen=True)
class NumericMode:
    name: str
    fmt: int          # _lib.FMT_BF16 / _lib.FMT_FP16
    a_split: bool     # activations stored as hi + lo planes
    w_split: bool     # weights packed as hi + lo
    raw_f32: bool     # raw conv outputs (pre-norm) in fp32
    # Mixed modes refine the three switches per buffer / per layer (None = the switch above applies everywhere):
    #   a_split_bufs    names of the engine's activation buffers stored as hi + lo ("in", "mid0", "cat0", "pool1", ...)
    #   w_split_layers  conv tags whose weights are split; raw_f32_layers conv tags whose raw output stays fp32.
    # Tags: "enc<l>.1|2" / "dec<l>.1|2" = conv1 | conv2 of the encoder / decoder ConvBlock3D at resolution level l
    # (0 = full resolution; "enc0" selects both convs), "up" = every ConvTranspose / 1x1 projection.
    a_split_bufs: Optional[frozenset] = None
    w_split_layers: Optional[frozenset] = None
    raw_f32_layers: Optional[frozenset] = None

    @property
    def passes(self) -> int:
        return 1 + int(self.a_split) + int(self.w_split)

    @property
    def mixed(self) -> bool:
        return self.a_split_bufs is not None or self.w_split_layers is not None or self.raw_f32_layers is not None

    def buffer(self, name: str) -> "NumericMode":
        """The mode an activation buffer is allocated in (decides hi-only vs hi + lo storage)."""
        if self.a_split_bufs is None:
            return self
        return _resolved(self, name in self.a_split_bufs, self.w_split, self.raw_f32)

    def layer(self, tag: str, src_split: Optional[bool] = None) -> "NumericMode":
        """The mode one conv layer runs in: its input buffer decides the activation split, the tag the rest."""
        if not self.mixed:
            return self
        a = self.a_split if src_split is None else bool(src_split)
        w = self.w_split if self.w_split_layers is None else _match(self.w_split_layers, tag)
        r = self.raw_f32 if self.raw_f32_layers is None else _match(self.raw_f32_layers, tag)
        return _resolved(self, a, w, r)

    @property
    def dtype(self) -> torch.dtype:
        return torch.float16 if self.fmt == _lib.FMT_FP16 else torch.bfloat16

    @property
    def conv_flags(self) -> int:
        return _lib.CONV_FP16_FLAG if self.fmt == _lib.FMT_FP16 else 0

    @property
    def bench_dtype(self) -> str:
        base = "fp16" if self.fmt == _lib.FMT_FP16 else "bf16"
        if self.mixed:
            return f"{base}-mixed"
        return base if self.passes == 1 else f"{base}x{self.passes}"


@lru_cache(maxsize=None)
def _resolved(nm: "NumericMode", a: bool, w: bool, r: bool) -> "NumericMode":
    return replace(nm, a_split=a, w_split=w, raw_f32=r, a_split_bufs=None, w_split_layers=None, raw_f32_layers=None)


def _tags(env: str, default: str) -> frozenset:
    return frozenset(t for t in os.environ.get(env, default).split(",") if t)


MODES = {
    "bf16": NumericMode("bf16", _lib.FMT_BF16, False, False, False),
    "fp16": NumericMode("fp16", _lib.FMT_FP16, False, False, False),
    "fp16w2": NumericMode("fp16w2", _lib.FMT_FP16, False, True, True),
    "fp16a2": NumericMode("fp16a2", _lib.FMT_FP16, True, False, True),
    "parity": NumericMode("parity", _lib.FMT_BF16, True, True, True),
    "fp16x3": NumericMode("fp16x3", _lib.FMT_FP16, True, True, True),
    # the gate-passing fast mode: fp16 single pass everywhere, except that the buffers / layers which dominate the
    # error carry more bits (see the module docstring); MMSEG_FP16M_A / _W override the sets for probing
    "fp16m": NumericMode("fp16m", _lib.FMT_FP16, False, False, True, _tags("MMSEG_FP16M_A", "in,mid0,pool1"),
                         _tags("MMSEG_FP16M_W", "enc0"), None),
}
# the leanest gate-passing rung found on B200 (tools/ladder_probe.py, 36-window crop): only the INPUT volume hi + lo and the
# first block's weights split, fp32 raw outputs — 9.1e-4 rel-L2 / 99.92 % labels / Dice 0.99917 (fp16m: 6.8e-4 / 99.94 %);
# the margins are thin, so the drop-in default stays fp16m and bench.py headlines whichever rung it MEASURES as passing
MODES["fp16i"] = NumericMode("fp16i", _lib.FMT_FP16, False, False, True, frozenset({"in"}), frozenset({"enc0"}), None)
MODES["bf16x3"] = MODES["parity"]

# fastest first: bench.py walks this ladder and headlines the first mode that meets every gate
LADDER = ("bf16", "fp16", "fp16i", "fp16m", "fp16w2", "fp16a2", "parity")
# inference default of the drop-in models: the fastest rung that meets every gate
DEFAULT_INFERENCE_MODE = "fp16m"


def mode(m: Union[str, bool, "NumericMode", None]) -> NumericMode:
    """Accepts a mode name, a NumericMode, or the round-1 `split` boolean (False = bf16, True = parity)."""
    if isinstance(m, NumericMode):
        return m
    if m is None or m is False:
        return MODES["bf16"]
    if m is True:
        return MODES["parity"]
    try:
        return MODES[m]
    except KeyError:
        raise ValueError(f"unknown numeric mode {m!r}; choose from {sorted(MODES)}") from None
