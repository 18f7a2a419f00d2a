"""Oracle restatement of ``monai.inferers.sliding_window_inference`` (CPU, torch).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

PARITY UNPINNED: the reference delegates this arithmetic to MONAI
(requirements.txt:7 ``monai>=1.3.0`` — unpinned, not vendored, not installed in
the build container; the reference has no tests/golden vectors).  This file
restates MONAI >= 1.3's published algorithm as specified in SURVEY.md
Appendix C and is anchored on the reference's only call site,
src/trainer/trainer.py:381-392:

    sliding_window_inference(image, roi_size=tuple(roi), sw_batch_size=B,
                             predictor=self.model, overlap=overlap)

i.e. mode="constant" unless the caller passes one; sigma_scale=0.125;
padding_mode="constant", cval=0.
"""
import math
from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def scan_interval(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[int]:
    """MONAI _get_scan_interval: roi if roi == image else max(int(roi*(1-overlap)), 1)."""
    out = []
    for im, r in zip(image_size, roi_size):
        if r == im:
            out.append(int(r))
        else:
            iv = int(r * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return out


def axis_starts(image: int, roi: int, interval: int) -> List[int]:
    """MONAI dense_patch_slices, one axis: first d with d*interval+roi >= image, starts clamped."""
    num = int(math.ceil(float(image) / interval))
    scan_dim = next((d for d in range(num) if d * interval + roi >= image), None)
    n = scan_dim + 1 if scan_dim is not None else 1
    starts = []
    for idx in range(n):
        s = idx * interval
        s -= max(s + roi - image, 0)
        starts.append(s)
    return starts


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[Tuple[int, ...]]:
    """All window origins, axis 0 slowest / last axis fastest (np.meshgrid(indexing='ij') order)."""
    iv = scan_interval(image_size, roi_size, overlap)
    per_axis = [axis_starts(im, r, i) for im, r, i in zip(image_size, roi_size, iv)]
    out: List[Tuple[int, ...]] = [()]
    for starts in per_axis:
        out = [o + (s,) for o in out for s in starts]
    return out


def importance_map(roi_size: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125,
                   dtype=torch.float32) -> Tensor:
    """MONAI compute_importance_map + the clamp at max(min, 1e-3) applied by sliding_window_inference."""
    if mode == "constant":
        w = torch.ones(tuple(roi_size), dtype=dtype)
    elif mode == "gaussian":
        w = None
        for i, r in enumerate(roi_size):
            sigma = r * sigma_scale
            x = torch.arange(-(r - 1) / 2.0, (r - 1) / 2.0 + 1, dtype=dtype)
            g = torch.exp(x ** 2 / (-2 * sigma ** 2))
            w = g if w is None else w.unsqueeze(-1) * g[(None,) * i]
    else:
        raise ValueError(mode)
    floor = max(float(w.min()), 1e-3)
    return w.clamp_(min=floor)


def sliding_window_inference(inputs: Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[Tensor], Tensor], overlap: float = 0.25,
                             mode: str = "constant", sigma_scale: float = 0.125, cval: float = 0.0,
                             return_count: bool = False):
    """Appendix C steps 1-6.  inputs [B, C, *spatial] -> [B, out_channels, *spatial] (input dtype)."""
    nsp = inputs.dim() - 2
    B = inputs.shape[0]
    orig = list(inputs.shape[2:])
    roi = [int(r) if r and r > 0 else o for r, o in zip(roi_size, orig)]
    image_size = [max(o, r) for o, r in zip(orig, roi)]
    pad = []
    for k in range(nsp - 1, -1, -1):
        diff = max(roi[k] - orig[k], 0)
        half = diff // 2
        pad.extend([half, diff - half])
    if any(pad):
        inputs = F.pad(inputs, pad, mode="constant", value=cval)
    starts = window_starts(image_size, roi, overlap)
    num_win = len(starts)
    w = importance_map(roi, mode, sigma_scale, inputs.dtype)
    output = None
    count = torch.zeros((1, 1, *image_size), dtype=inputs.dtype)
    total = num_win * B
    for g in range(0, total, sw_batch_size):
        idxs = range(g, min(g + sw_batch_size, total))
        sl = []
        for idx in idxs:
            b, wi = idx // num_win, idx % num_win
            sl.append((slice(b, b + 1), slice(None)) + tuple(slice(s, s + r) for s, r in zip(starts[wi], roi)))
        win = torch.cat([inputs[s] for s in sl])
        seg = predictor(win)
        if output is None:
            output = torch.zeros((B, seg.shape[1], *image_size), dtype=seg.dtype)
        seg = seg * w
        for j, s in enumerate(sl):
            output[s] += seg[j]
            if s[0].start == 0:  # count map is [1,1,*]: data-independent, accumulated once per window
                count[(slice(0, 1), slice(None)) + s[2:]] += w
    output = output / count
    if any(pad):
        crop = [slice(None), slice(None)]
        for k in range(nsp):
            before = pad[2 * (nsp - 1 - k)]
            crop.append(slice(before, before + orig[k]))
        output = output[tuple(crop)]
    if return_count:
        return output, count
    return output
