"""Device-side counterparts of the reference's inference-edge transforms (src/data/transforms.py)."""
from .transforms import ModalitySpecificNormalize, Resize

__all__ = ["ModalitySpecificNormalize", "Resize"]
