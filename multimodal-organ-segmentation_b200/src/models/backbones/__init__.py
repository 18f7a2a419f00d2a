"""Backbones with sm_100a kernels behind the reference's class names (UNet3D, DualEncoder).  SwinUNETR — a thin wrapper
over MONAI in the reference — is scope row N2 and not built."""
from . import dual_encoder as _de, unet as _unet

UNet3D = _unet.UNet3D
DualEncoder = _de.DualEncoder

__all__ = ("UNet3D", "DualEncoder")
