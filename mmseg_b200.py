"""Importable name for the package directory `multimodal-organ-segmentation_b200/` (not a Python identifier).

`import mmseg_b200` loads that directory as the package `mmseg_b200`, so submodules are `mmseg_b200.engine`,
`mmseg_b200.src.models`, ... with a single module identity.
"""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_root, "multimodal-organ-segmentation_b200")
_spec = importlib.util.spec_from_file_location(
    "mmseg_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["mmseg_b200"] = _pkg
_spec.loader.exec_module(_pkg)
