"""Backbones with sm_100a kernels behind the reference's class names (SwinUNETR, UNet3D, DualEncoder; reference
src/models/backbones/__init__.py:5-13)."""
from . import dual_encoder as _de, swin_unetr as _swin, unet as _unet

SwinUNETR = _swin.SwinUNETR
UNet3D = _unet.UNet3D
DualEncoder = _de.DualEncoder

__all__ = ("SwinUNETR", "UNet3D", "DualEncoder")
