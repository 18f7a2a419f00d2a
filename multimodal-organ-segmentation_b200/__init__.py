"""multimodal-organ-segmentation_b200 — B200 (sm_100a) drop-in for the reference's segmentation hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/mmseg_b200.h) -> libmmseg_b200.so
  _lib.py          ctypes binding (raises if the library is missing; no CPU fallback)
  kernels.py       torch-facing wrappers (device memory + stream plumbing only)
  tiling.py        host-side tile planner for the tcgen05 conv kernel
  engine.py        launch sequences for UNet3D / DualEncoder forward
  src/             mirror of the reference's `src` package for this path (same names, signatures, state_dict)

The directory name is not a Python identifier; import it with
    importlib.import_module("multimodal-organ-segmentation_b200")
or through the `mmseg_b200` alias module at the repository root.
"""
__version__ = "0.1.0"
