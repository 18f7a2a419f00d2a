"""Model factory — drop-in mirror of the reference's src/models/build.py (registry, wrapper, checkpoint I/O)."""
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .backbones.unet import UNet3D, build_unet3d
from .backbones.dual_encoder import DualEncoder, build_dual_encoder


def _build_swin_unetr(config: Dict[str, Any]) -> nn.Module:
    # reference build.py:17 -> backbones/swin_unetr.py:71-96 wraps monai.networks.nets.SwinUNETR (third-party, absent)
    raise NotImplementedError("swin_unetr is scope row N2 (MONAI SwinUNETR restatement): not part of this build")


# reference build.py:16-21
MODEL_REGISTRY = {
    "swin_unetr": _build_swin_unetr,
    "unet": build_unet3d,
    "unet3d": build_unet3d,
    "dual_encoder": build_dual_encoder,
}


class MultiModalSegmentationModel(nn.Module):
    """reference build.py:24-74: thin wrapper, state_dict prefix `backbone.`."""

    def __init__(self, backbone: nn.Module, config: Dict[str, Any]):
        super().__init__()
        self.backbone = backbone
        self.config = config
        self.num_modalities = len(config["data"]["modalities"])

    def forward(self, x: torch.Tensor, return_features: bool = False):
        return self.backbone(x, return_features=return_features)

    def set_numeric_mode(self, mode: str) -> "MultiModalSegmentationModel":
        self.backbone.set_numeric_mode(mode)
        return self

    def load_pretrained(self, path: str) -> None:
        if hasattr(self.backbone, "load_pretrained"):
            self.backbone.load_pretrained(path)
        else:
            state_dict = torch.load(path, map_location="cpu")
            if "model_state_dict" in state_dict:
                state_dict = state_dict["model_state_dict"]
            self.load_state_dict(state_dict, strict=False)


def build_model(config: Dict[str, Any]) -> nn.Module:
    """reference build.py:77-114 (including the in_channels mutation at :98-99 and the device move at :108-112)."""
    model_name = config["model"]["name"].lower()
    if model_name not in MODEL_REGISTRY:
        raise ValueError(f"Unknown model: {model_name}. Available: {list(MODEL_REGISTRY.keys())}")
    num_modalities = len(config["data"]["modalities"])
    if model_name in ["swin_unetr", "unet", "unet3d"]:
        config["model"]["in_channels"] = num_modalities
    backbone = MODEL_REGISTRY[model_name](config)
    model = MultiModalSegmentationModel(backbone, config)
    mode = config.get("hardware", {}).get("numeric_mode")  # optional new key; default = bf16 throughput mode
    if mode:
        model.set_numeric_mode(mode)
    device = config["hardware"]["device"]
    if device == "cuda" and torch.cuda.is_available():
        model = model.cuda()
    return model


def get_model(config: Dict[str, Any]) -> nn.Module:
    return build_model(config)


def load_checkpoint(model: nn.Module, checkpoint_path: str, strict: bool = False) -> Dict[str, Any]:
    """reference build.py:122-150."""
    checkpoint = torch.load(checkpoint_path, map_location="cpu")
    if "model_state_dict" in checkpoint:
        state_dict = checkpoint["model_state_dict"]
    elif "state_dict" in checkpoint:
        state_dict = checkpoint["state_dict"]
    else:
        state_dict = checkpoint
    model.load_state_dict(state_dict, strict=strict)
    return checkpoint


def save_checkpoint(model: nn.Module, optimizer: Optional[torch.optim.Optimizer], epoch: int, checkpoint_path: str,
                    **kwargs) -> None:
    """reference build.py:153-180."""
    checkpoint = {"epoch": epoch, "model_state_dict": model.state_dict()}
    if optimizer is not None:
        checkpoint["optimizer_state_dict"] = optimizer.state_dict()
    checkpoint.update(kwargs)
    torch.save(checkpoint, checkpoint_path)
