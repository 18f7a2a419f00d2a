"""Mirror of reference src/models/backbones/__init__.py (SwinUNETR is scope row N2: not built)."""
from .unet import UNet3D
from .dual_encoder import DualEncoder

__all__ = ["UNet3D", "DualEncoder"]
