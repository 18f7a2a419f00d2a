"""Install the UNMODIFIED reference modules of the hot path into baseline/_ref/ (git-ignored, NOT gpurun-ignored, so the
copy travels to the GPU box like a built .so) — the reference arm of bench.py and the N=1 `cpu_baseline` then time the
reference's own `build_model` / `forward` instead of the oracle port.

The reference has no setup.py / pyproject.toml (nothing for `pip install --target baseline/_ref` to build), so the
install is a plain copy of the packages the path needs: src/__init__.py, src/models/**, src/trainer/** and — for the
drop-in test of the CLI — main.py, src/utils/**, src/data/**, configs/default.yaml (stock files, byte for byte).  Nothing under baseline/_ref is ever committed, imported by the product, or edited.

TEST / BENCH INFRASTRUCTURE ONLY: tests/, __graft_entry__ and bench.py's cpu legs may import it; the product never does.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
# the hot-path packages + the CLI and its host-side helpers (tests/test_dropin_main.py drives the reference's own
# main.py with src.models / src.trainer shadowed by this repo's mirror)
PARTS = ["src/__init__.py", "src/models", "src/trainer", "src/utils", "src/data", "main.py", "configs/default.yaml"]


def install(force: bool = False) -> bool:
    """Copy the reference packages when /root/reference is present (build container); returns True when baseline/_ref
    is usable afterwards (on the GPU box only the shipped copy exists)."""
    if os.path.isdir(os.path.join(REF, "src", "models")):
        for part in PARTS:
            src, dst = os.path.join(REF, part), os.path.join(DST, part)
            if os.path.isdir(src):
                if force and os.path.isdir(dst):
                    shutil.rmtree(dst)
                shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            else:
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
    return available()


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "models", "build.py"))


def import_reference():
    """(build_model, get_loss) of the stock reference, imported from baseline/_ref as the top-level package `src`."""
    if not available():
        raise ImportError("baseline/_ref is missing: run `python oracle/install_ref.py` in the build container")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from src.models import build_model          # noqa: E402  (the reference's own factory)
    from src.trainer.losses import get_loss     # noqa: E402
    return build_model, get_loss


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print(f"baseline/_ref {'ready' if ok else 'NOT available'} at {DST}")
