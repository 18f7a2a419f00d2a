"""Model package of the sm_100a path: the factory entry points the reference's main.py imports
(`build_model`, `get_model`; reference src/models/__init__.py:5)."""
from . import build as _build

build_model = _build.build_model
get_model = _build.get_model

__all__ = ("build_model", "get_model")
