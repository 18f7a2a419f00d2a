"""ModalitySpecificNormalize / Resize of the reference's data pipeline (src/data/transforms.py:362-404, 215-250) on the
device — SURVEY.md §8(f) N3: at ~0.5 s of GPU time per 512x512x300 volume the three numpy sweeps of the host version
would dominate an inference case, so the volume is uploaded raw and normalised in HBM by two streaming kernels.

Same constructor arguments and the same sample-dict call convention as the reference classes, but sample["image"] is a
CUDA fp32 tensor [C, H, W, D].  There is no CPU route (a host tensor raises): the host implementation IS the reference.
Only what the inference edge needs is built: images (not labels; the reference resizes labels with nearest neighbour
for training augmentation, which stays on the host with the rest of the augmentation pipeline).
"""
import ctypes as C
from typing import Any, Dict, Tuple

import torch

from ... import kernels as K

_KIND = {"CT": 1, "PET": 2, "MRI": 3, "US": 3}


def _require_cuda_image(sample: Dict[str, Any]) -> torch.Tensor:
    image = sample["image"]
    if not (torch.is_tensor(image) and image.is_cuda):
        raise RuntimeError("mmseg_b200 device transforms take sample['image'] as a CUDA tensor [C, H, W, D]; the host "
                           "(numpy) implementation is the reference's own src/data/transforms.py")
    if image.dim() != 4:
        raise ValueError("image must be [C, H, W, D]")
    return image.contiguous().float()


class ModalitySpecificNormalize:
    """CT: window clip + rescale to [0, 1]; PET: divide by the volume maximum; MRI / US: z-score — per channel in the
    order of config['data']['modalities'] (reference transforms.py:362-404)."""

    def __init__(self, config: Dict[str, Any]):
        self.config = config
        self.modalities = config["data"]["modalities"]
        self.preprocess_config = config["data"]["preprocessing"]

    def _tables(self, n_channels: int, device):
        kind, a, b = [0] * n_channels, [0.0] * n_channels, [1.0] * n_channels
        for c, modality in enumerate(self.modalities[:n_channels]):
            mc = self.preprocess_config.get(modality.lower(), {})
            if modality == "CT":
                center, width = mc.get("window_center", 0), mc.get("window_width", 400)
                kind[c], a[c], b[c] = 1, center - width / 2, center + width / 2
            elif modality == "PET":
                kind[c] = 2 if mc.get("normalize", True) else 0
            elif modality in ("MRI", "US"):
                kind[c] = 3 if mc.get("normalize", True) else 0
        return (torch.tensor(kind, dtype=torch.int32, device=device), torch.tensor(a, dtype=torch.float32, device=device),
                torch.tensor(b, dtype=torch.float32, device=device))

    def __call__(self, sample: Dict[str, Any]) -> Dict[str, Any]:
        image = _require_cuda_image(sample)
        Cc = image.shape[0]
        nvox = image[0].numel()
        key = (Cc, str(image.device))
        if getattr(self, "_key", None) != key:
            self._tab, self._key = self._tables(Cc, image.device), key
        kind, a, b = self._tab
        n_blocks = max(1, min(148 * 4, (nvox + 65535) // 65536))
        partial = torch.empty((Cc, n_blocks, 3), dtype=torch.float64, device=image.device)
        stats = torch.empty((Cc, 3), dtype=torch.float32, device=image.device)
        out = torch.empty_like(image)
        K._call("mmseg_channel_stats", K._ptr(image), Cc, nvox, K._ptr(partial), n_blocks, K._ptr(stats), K._stream())
        K._call("mmseg_modality_normalize", K._ptr(image), K._ptr(out), Cc, nvox, K._ptr(kind), K._ptr(a), K._ptr(b),
                K._ptr(stats), K._stream())
        sample["image"] = out
        return sample


class Resize:
    """scipy.ndimage.zoom(order=1) of every channel to `size` == trilinear interpolation with aligned corners
    (reference transforms.py:215-250)."""

    def __init__(self, size: Tuple[int, int, int], order: int = 1):
        if order != 1:
            raise NotImplementedError("device Resize implements order=1 (the reference's default); other orders stay on the host")
        self.size, self.order = tuple(int(s) for s in size), order

    def __call__(self, sample: Dict[str, Any]) -> Dict[str, Any]:
        image = _require_cuda_image(sample)
        if "label" in sample:
            raise NotImplementedError("label resizing (nearest neighbour, training augmentation) stays on the host")
        sample["image"] = K.trilinear_resize(image.unsqueeze(0), self.size)[0]
        return sample
