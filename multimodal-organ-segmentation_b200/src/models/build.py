"""Model factory for the sm_100a path.

API-compatible with the reference's src/models/build.py (same public names, registry keys, config keys read, wrapper
attribute `backbone`, checkpoint dictionary keys), so `main.py` and existing checkpoints work unchanged; written
independently around a small table of builders.
"""
from typing import Any, Callable, Dict, Optional

import torch
import torch.nn as nn

from .backbones.dual_encoder import DualEncoder, build_dual_encoder  # noqa: F401  (re-exported like the reference)
from .backbones.swin_unetr import SwinUNETR, build_swin_unetr        # noqa: F401
from .backbones.unet import UNet3D, build_unet3d                      # noqa: F401

Config = Dict[str, Any]


# name -> builder(config); the keys are the reference's MODEL_REGISTRY keys (build.py:16-21)
MODEL_REGISTRY: Dict[str, Callable[[Config], nn.Module]] = {
    "swin_unetr": build_swin_unetr,
    "unet": build_unet3d,
    "unet3d": build_unet3d,
    "dual_encoder": build_dual_encoder,
}
# early-fusion backbones take one input channel per modality (reference build.py:98-99 rewrites the config in place)
_CHANNELS_FROM_MODALITIES = ("swin_unetr", "unet", "unet3d")


class MultiModalSegmentationModel(nn.Module):
    """Wrapper around the backbone: its parameters live under the `backbone.` state_dict prefix (reference build.py:24-74)."""

    def __init__(self, backbone: nn.Module, config: Config):
        super().__init__()
        self.backbone, self.config = backbone, config
        self.num_modalities = len(config["data"]["modalities"])

    def forward(self, x: torch.Tensor, return_features: bool = False):
        return self.backbone(x, return_features=return_features)

    def set_numeric_mode(self, mode: str) -> "MultiModalSegmentationModel":
        """'bf16' (throughput) | 'parity' (3-pass split bf16) — an option of this path, not of the reference."""
        self.backbone.set_numeric_mode(mode)
        return self

    def load_pretrained(self, path: str) -> None:
        loader = getattr(self.backbone, "load_pretrained", None)
        if callable(loader):
            loader(path)
            return
        blob = torch.load(path, map_location="cpu")
        self.load_state_dict(blob.get("model_state_dict", blob), strict=False)


def build_model(config: Config) -> nn.Module:
    """config["model"]["name"] -> wrapped backbone on config["hardware"]["device"] (reference build.py:77-114)."""
    name = str(config["model"]["name"]).lower()
    builder = MODEL_REGISTRY.get(name)
    if builder is None:
        raise ValueError(f"Unknown model: {name}. Available: {list(MODEL_REGISTRY.keys())}")
    if name in _CHANNELS_FROM_MODALITIES:
        config["model"]["in_channels"] = len(config["data"]["modalities"])
    model = MultiModalSegmentationModel(builder(config), config)
    numeric_mode = config.get("hardware", {}).get("numeric_mode")   # optional key of this path; default bf16
    if numeric_mode:
        model.set_numeric_mode(numeric_mode)
    if config["hardware"]["device"] == "cuda" and torch.cuda.is_available():
        model = model.cuda()
    return model


def get_model(config: Config) -> nn.Module:
    """Alias kept for the reference's `from src.models import get_model`."""
    return build_model(config)


def _state_dict_of(blob: Dict[str, Any]) -> Dict[str, Any]:
    for key in ("model_state_dict", "state_dict"):
        if key in blob:
            return blob[key]
    return blob


def load_checkpoint(model: nn.Module, checkpoint_path: str, strict: bool = False) -> Dict[str, Any]:
    """Loads the weights and returns the whole checkpoint dictionary (epoch, optimizer state, ... — build.py:122-150)."""
    blob = torch.load(checkpoint_path, map_location="cpu")
    model.load_state_dict(_state_dict_of(blob), strict=strict)
    return blob


def save_checkpoint(model: nn.Module, optimizer: Optional[torch.optim.Optimizer], epoch: int, checkpoint_path: str,
                    **extra) -> None:
    """{"epoch", "model_state_dict"[, "optimizer_state_dict"], **extra} — the layout main.py:361,398 reads back."""
    blob: Dict[str, Any] = {"epoch": epoch, "model_state_dict": model.state_dict()}
    if optimizer is not None:
        blob["optimizer_state_dict"] = optimizer.state_dict()
    blob.update(extra)
    torch.save(blob, checkpoint_path)
