"""The C-ABI library loads and exports every symbol include/mmseg_b200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

import mmseg_b200  # noqa: F401
from mmseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mmseg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmseg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmseg_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in _lib.SYMBOLS"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"


def test_version_and_error_channel():
    assert _lib.lib.mmseg_version() == 2
    # a rejected call returns a negative status and sets the thread-local message; nothing is launched
    a = _lib.ConvArgs()
    a.ksize = 5
    rc = _lib.lib.mmseg_conv3d_smem_bytes(ctypes.byref(a))
    assert rc == -2 and "ksize" in _lib.last_error()


def test_product_fails_loudly_without_device():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from mmseg_b200.src.models.backbones.unet import UNet3D
    m = UNet3D(2, 8, [16, 32])
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 16, 16, 16))
    with pytest.raises(RuntimeError):
        _lib.require_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodal-organ-segmentation_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(d, f)


def test_struct_layouts_match_the_library():
    """Every argument struct of the ABI has the same size in the ctypes binding and in the compiled library."""
    structs = (_lib.ConvArgs, _lib.WgradArgs, _lib.NormArgs, _lib.NormBwdArgs, _lib.AdamwTensor, _lib.RepackDesc, _lib.SwinAttnArgs)
    for which, st in enumerate(structs):
        assert _lib.lib.mmseg_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert _lib.lib.mmseg_sizeof(len(structs)) == -1
