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


@pytest.mark.parametrize("name,fusion,norm", [("unet", "early", "instance"), ("unet", "early", "batch"),
                                              ("dual_encoder", "attention", "group"), ("dual_encoder", "late", "batch"),
                                              ("dual_encoder", "concat", "none")])
def test_oracle_training_step_matches_reference_autograd(ref, name, fusion, norm):
    """oracle/train.py (the checker of the kernel training path) against the reference's own train-mode modules + DiceCELoss
    under autograd in fp64: loss and every parameter gradient, for the norm kinds of model.backbone.norm."""
    from oracle.train import train_step
    torch.manual_seed(21)
    cfg = _cfg(name, fusion, ["CT", "PET"], [8, 16])
    cfg["model"]["backbone"]["norm"] = norm
    m = ref["build"].build_model(cfg).double().train()
    with torch.no_grad():
        for key, p in m.named_parameters():
            if ".norm" in key:
                p.uniform_(0.6, 1.4) if key.endswith("weight") else p.normal_(0, 0.3)
    x = torch.randn(2, 2, 16, 16, 16, dtype=torch.float64)
    y = torch.randint(0, 8, (2, 16, 16, 16))
    sd = {k[len("backbone."):]: v.detach().clone() for k, v in m.state_dict().items()}
    crit = ref["losses"].DiceCELoss()
    loss = crit(m(x), y)
    loss.backward()
    kind = "unet" if name == "unet" else "dual"
    got_loss, got_g, _ = train_step(kind, sd, dict(L=2, M=2, fusion=fusion, norm=norm), x, y)
    assert abs(got_loss - loss.item()) < 1e-9 * max(1.0, abs(loss.item()))
    scale = max(p.grad.norm().item() for p in m.parameters())
    for key, p in m.named_parameters():
        g = got_g[key[len("backbone."):]]
        assert g is not None, key
        assert (g - p.grad).norm().item() < 1e-8 * scale, (key, (g - p.grad).norm().item(), scale)


def test_oracle_training_step_dropout_mask_matches_reference(ref):
    """Dropout3d before out_conv (unet.py:162,198): the oracle step takes the [n, C] mask / (1 - p) the reference drew."""
    from oracle.train import train_step
    torch.manual_seed(22)
    cfg = _cfg("unet", "early", ["CT", "PET"], [8, 16])
    cfg["model"]["head"]["dropout"] = 0.5
    m = ref["build"].build_model(cfg).double().train()
    drawn = {}

    def hook(_mod, inp, out):
        a, b = inp[0].detach(), out.detach()
        drawn["drop"] = (b.flatten(2).abs().sum(-1) / a.flatten(2).abs().sum(-1).clamp_min(1e-300))
    h = m.backbone.dropout.register_forward_hook(hook)
    x = torch.randn(2, 2, 16, 16, 16, dtype=torch.float64)
    y = torch.randint(0, 8, (2, 16, 16, 16))
    loss = ref["losses"].DiceCELoss()(m(x), y)
    loss.backward()
    h.remove()
    drop = drawn["drop"]
    assert set(drop.flatten().round(decimals=6).tolist()) <= {0.0, 2.0} and (drop == 0).any() and (drop > 0).any()
    sd = {k[len("backbone."):]: v.detach().clone() for k, v in m.state_dict().items()}
    got_loss, got_g, _ = train_step("unet", sd, dict(L=2, norm="instance", drop=drop), x, y)
    assert abs(got_loss - loss.item()) < 1e-9
    scale = max(p.grad.norm().item() for p in m.parameters())
    for key, p in m.named_parameters():
        assert (got_g[key[len("backbone."):]] - p.grad).norm().item() < 1e-8 * scale, key
