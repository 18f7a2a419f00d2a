"""The oracle against the reference's own modules imported read-only (build container only; skipped on the GPU box)."""
import copy
import sys

import pytest
import torch

from oracle import losses as OL
from oracle import models as OM
from oracle import sliding_window as OS


@pytest.fixture(scope="module")
def ref(reference_root):
    sys.path.insert(0, reference_root)
    try:
        import src.models.build as build
        import src.trainer.losses as losses
        yield {"build": build, "losses": losses}
    finally:
        sys.path.remove(reference_root)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]


def _cfg(name, fusion, mods, features):
    return {"model": {"name": name, "in_channels": len(mods), "out_channels": 8,
                      "backbone": {"features": features}, "fusion": {"type": fusion}, "head": {"dropout": 0.0}},
            "data": {"modalities": mods}, "hardware": {"device": "cpu"}}


def test_unet_default_features_matches_reference(ref):
    """Full-width UNet3D (features 32..512) on a 32^3 input, fresh seed: oracle == reference module."""
    torch.manual_seed(11)
    m = ref["build"].build_model(_cfg("unet", "early", ["CT", "PET"], [32, 64, 128, 256, 512])).eval()
    x = torch.randn(1, 2, 32, 32, 32)
    with torch.no_grad():
        want = m(x)
    got = OM.unet3d_forward(m.state_dict(), x)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("fusion", ["late", "attention", "concat"])
def test_dual_encoder_matches_reference(ref, fusion):
    torch.manual_seed(12)
    m = ref["build"].build_model(_cfg("dual_encoder", fusion, ["CT", "PET"], [16, 32, 64])).eval()
    x = torch.randn(1, 2, 16, 16, 16)
    with torch.no_grad():
        want = m(x)
    assert torch.allclose(OM.dual_encoder_forward(m.state_dict(), x, fusion), want, atol=2e-5, rtol=1e-5)


def test_dicece_matches_reference(ref):
    torch.manual_seed(13)
    lg, tg = torch.randn(2, 8, 9, 7, 11), torch.randint(0, 8, (2, 9, 7, 11))
    want = ref["losses"].DiceCELoss()(lg, tg).item()
    assert abs(OL.dice_ce_loss(lg, tg)[0].item() - want) < 1e-6


def test_sliding_window_single_window_equals_forward(ref):
    """Appendix C invariant with the REFERENCE model as predictor: a roi-sized volume is one plain forward."""
    torch.manual_seed(14)
    m = ref["build"].build_model(_cfg("unet", "early", ["CT", "PET"], [16, 32])).eval()
    x = torch.randn(1, 2, 16, 16, 16)
    with torch.no_grad():
        want = m(x)
        got = OS.sliding_window_inference(x, (16, 16, 16), 4, m, overlap=0.5, mode="gaussian")
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
