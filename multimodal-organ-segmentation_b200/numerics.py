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
  fp16m    fp16      2 / 3   fp16a2 on the big layers, fp16x3 on the cheap ones (per-layer policy, below)

Measured on B200 (36-window bench crop vs the fp32 reference): with exact activations the residual error of fp16a2 is
the weights' fp16 rounding alone — 1.07e-3 relative L2, 7 % over the gate — while exact weights (fp16w2) barely help
(2.7e-3): a static weight perturbation is partly removed by the InstanceNorm that follows (its per-channel mean over
non-negative inputs is constant across voxels), a per-voxel activation perturbation is not.  `fp16m` therefore splits
the activations everywhere and the weights only where the third pass is cheap: the ConvTranspose GEMMs (no norm behind
them) and the levels at or below `w_split_min_level` (1/8, 1/64, ... of the voxels per level).

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


@dataclass(frozen=True)
class NumericMode:
    name: str
    fmt: int          # _lib.FMT_BF16 / _lib.FMT_FP16
    a_split: bool     # activations stored as hi + lo planes
    w_split: bool     # weights packed as hi + lo
    raw_f32: bool     # raw conv outputs (pre-norm) in fp32
    # mixed modes: layers at resolution level >= w_split_min_level (0 = full resolution) and every ConvTranspose also
    # split their weights (3 passes); None = `w_split` applies to every layer alike
    w_split_min_level: Optional[int] = None

    @property
    def passes(self) -> int:
        return 1 + int(self.a_split) + int(self.w_split)

    def for_layer(self, level: int, is_convt: bool = False) -> "NumericMode":
        """The mode one conv layer runs in (same element format and activation layout; only the weight split varies)."""
        if self.w_split_min_level is None or self.w_split:
            return self
        if is_convt or level >= self.w_split_min_level:
            return _with_w_split(self)
        return self

    @property
    def dtype(self) -> torch.dtype:
        return torch.float16 if self.fmt == _lib.FMT_FP16 else torch.bfloat16

    @property
    def conv_flags(self) -> int:
        return _lib.CONV_FP16_FLAG if self.fmt == _lib.FMT_FP16 else 0

    @property
    def bench_dtype(self) -> str:
        base = "fp16" if self.fmt == _lib.FMT_FP16 else "bf16"
        if self.w_split_min_level is not None and not self.w_split:
            return f"{base}x{self.passes}-{self.passes + 1}"
        return base if self.passes == 1 else f"{base}x{self.passes}"


@lru_cache(maxsize=None)
def _with_w_split(nm: "NumericMode") -> "NumericMode":
    return replace(nm, w_split=True, w_split_min_level=None)


MODES = {
    "bf16": NumericMode("bf16", _lib.FMT_BF16, False, False, False),
    "fp16": NumericMode("fp16", _lib.FMT_FP16, False, False, False),
    "fp16w2": NumericMode("fp16w2", _lib.FMT_FP16, False, True, True),
    "fp16a2": NumericMode("fp16a2", _lib.FMT_FP16, True, False, True),
    "parity": NumericMode("parity", _lib.FMT_BF16, True, True, True),
    "fp16x3": NumericMode("fp16x3", _lib.FMT_FP16, True, True, True),
    # activations split everywhere; weights split on the ConvTranspose GEMMs and from level MMSEG_FP16M_LEVEL down
    "fp16m": NumericMode("fp16m", _lib.FMT_FP16, True, False, True, int(os.environ.get("MMSEG_FP16M_LEVEL", "2"))),
}
MODES["bf16x3"] = MODES["parity"]

# fastest first: bench.py walks this ladder and headlines the first mode that meets every gate
LADDER = ("bf16", "fp16", "fp16w2", "fp16a2", "fp16m", "parity")


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
