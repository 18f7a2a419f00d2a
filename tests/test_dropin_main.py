"""The reference's CLI on top of the drop-in packages (CPU part): parse_args -> load_config -> merge_config_with_args ->
run_inference / run_train reach THIS repo's build_model and Trainer with the reference's own configuration — and stop
exactly where a B200 is required (there is no CPU fallback).  The GPU part (tests/test_gpu_parity.py) runs
`main.run_inference` end to end."""
import sys

import pytest
import torch

from tests import dropin

pytestmark = pytest.mark.skipif(not dropin.available(), reason="baseline/_ref (reference CLI) not installed: "
                                                              "run python oracle/install_ref.py in the build container")


@pytest.fixture()
def refmain():
    mod, restore = dropin.load_reference_main()
    try:
        yield mod
    finally:
        restore()


def _config(refmain, tmp_path, argv):
    import os
    old = sys.argv
    sys.argv = ["main.py"] + argv
    try:
        args = refmain.parse_args()
    finally:
        sys.argv = old
    config = refmain.load_config(os.path.join(dropin.REF, "configs", "default.yaml"))
    return refmain.merge_config_with_args(config, args), args


def test_main_imports_bind_the_dropin_packages(refmain):
    import mmseg_b200.src.models as our_models
    import mmseg_b200.src.trainer as our_trainer
    # what main.run_* import lazily
    from src.models import build_model
    from src.trainer import Trainer
    assert build_model is our_models.build_model and Trainer is our_trainer.Trainer
    # and the rest of the CLI is the reference's own code
    assert refmain.load_config.__module__ == "src.utils.io" and "baseline/_ref" in sys.modules["src.utils"].__file__


def test_reference_default_config_builds_the_dropin_models(refmain, tmp_path):
    """Every model/fusion the reference CLI offers (except swin_unetr, whose arithmetic lives in MONAI) is constructed by
    the drop-in factory from the reference's default.yaml + CLI overrides, with the reference's parameter names."""
    for model, fusion in (("unet", "early"), ("dual_encoder", "cross_attention"), ("dual_encoder", "attention"),
                          ("dual_encoder", "late")):
        config, _ = _config(refmain, tmp_path, ["--mode", "train", "--model", model, "--fusion", fusion, "--device", "cpu",
                                                "--modalities", "CT", "PET"])
        from src.models import build_model
        m = build_model(config)
        keys = list(m.state_dict().keys())
        assert all(k.startswith("backbone.") for k in keys)
        assert config["model"]["in_channels"] == 2 or model == "dual_encoder"
        if model == "unet":
            assert m.state_dict()["backbone.init_conv.conv1.weight"].shape == (32, 2, 3, 3, 3)
        else:
            assert m.state_dict()["backbone.encoders.0.init_conv.conv1.weight"].shape[1] == 1


def test_run_inference_reaches_the_dropin_trainer(refmain, tmp_path):
    """CPU box: the reference's run_inference loads the checkpoint into the drop-in model and constructs the drop-in
    Trainer, which refuses to run without a B200 — loudly, not through a CPU fallback."""
    from src.models import build_model
    ck = tmp_path / "ck.pth"
    config, _ = _config(refmain, tmp_path, ["--mode", "inference", "--model", "unet", "--checkpoint", str(ck), "--input",
                                            str(tmp_path), "--output", str(tmp_path / "pred"), "--output-dir", str(tmp_path),
                                            "--device", "cuda"])
    torch.manual_seed(0)
    torch.save({"model_state_dict": build_model(config).state_dict()}, ck)
    logger = refmain.setup_logger("dropin", log_dir=str(tmp_path)) if "log_dir" in refmain.setup_logger.__code__.co_varnames \
        else refmain.get_logger("dropin")
    if torch.cuda.is_available():
        pytest.skip("GPU present: the end-to-end run is covered by the gpu-marked test")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        refmain.run_inference(config, logger)
