"""Mirror of reference src/models/__init__.py:5."""
from .build import build_model, get_model

__all__ = ["build_model", "get_model"]
