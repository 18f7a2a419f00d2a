"""Helpers that run the REFERENCE's own CLI (main.py, installed unmodified under baseline/_ref by oracle/install_ref.py)
on top of this repo's drop-in packages — INTEGRATION.md §2: `src.models` and `src.trainer` are shadowed in sys.modules by
`mmseg_b200.src.models` / `mmseg_b200.src.trainer`; everything else main.py imports (src.utils, src.data) is the
reference's own code.  nibabel / matplotlib are not installed anywhere here, so they are stubbed: the stub nibabel
reads / writes .npy payloads behind the .nii file names main.py and Trainer.predict use."""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "main.py")) and os.path.isdir(os.path.join(REF, "src", "utils"))


def _stub_nibabel():
    nib = types.ModuleType("nibabel")

    class Nifti1Image:
        def __init__(self, data, affine, header=None):
            self._data, self.affine, self.header = np.asarray(data), affine, header

        def get_fdata(self):
            return self._data.astype(np.float64)

    def load(path):
        with open(path, "rb") as f:
            blob = np.load(f, allow_pickle=False)
            data, affine = np.array(blob["data"]), np.array(blob["affine"])
        return Nifti1Image(data, affine)

    def save(img, path):
        with open(path, "wb") as f:
            np.savez(f, data=img._data, affine=np.asarray(img.affine))

    nib.Nifti1Image, nib.load, nib.save = Nifti1Image, load, save
    nib.Nifti1Header = type("Nifti1Header", (), {})
    return nib


def write_volume(path, data, affine=None):
    """A file the stub nibabel can load (any extension: .nii / .nii.gz names are kept for the directory scan)."""
    with open(path, "wb") as f:
        np.savez(f, data=np.asarray(data), affine=np.eye(4) if affine is None else affine)


def read_volume(path):
    with open(path, "rb") as f:
        return np.array(np.load(f, allow_pickle=False)["data"])


def load_reference_main():
    """Imports baseline/_ref/main.py as module `refmain` with the shadowing in place; returns (module, restore())."""
    import mmseg_b200  # noqa: F401
    import mmseg_b200.src.models as our_models
    import mmseg_b200.src.trainer as our_trainer
    saved = {k: sys.modules.get(k) for k in ("nibabel", "matplotlib", "matplotlib.pyplot", "src", "src.models", "src.trainer",
                                            "src.utils", "src.data", "refmain")}
    saved_path = list(sys.path)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        saved.setdefault(k, sys.modules[k])
        del sys.modules[k]
    sys.modules["nibabel"] = _stub_nibabel()
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.path.insert(0, REF)
    importlib.import_module("src")                 # the reference's top-level package (utils, data come from it)
    sys.modules["src.models"] = our_models         # ... but models and trainer are THIS repo's drop-ins
    sys.modules["src.trainer"] = our_trainer
    spec = importlib.util.spec_from_file_location("refmain", os.path.join(REF, "main.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refmain"] = mod
    spec.loader.exec_module(mod)

    def restore():
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, restore
