"""The oracle (oracle/) against the golden vectors generated from the reference itself (tests/golden/make_golden.py).

CPU only.  This is what pins the oracle on machines where /root/reference does not exist.
"""
import os

import pytest
import torch

from oracle import losses as OL
from oracle import models as OM

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def _close(a, b, atol=2e-5, rtol=1e-5):
    assert a.shape == b.shape
    assert torch.allclose(a, b, atol=atol, rtol=rtol), (a - b).abs().max().item()


def test_unet_small_logits():
    g = load("unet_small")
    y, feats = OM.unet3d_forward(g["state_dict"], g["x"], return_features=True)
    _close(y, g["logits"])
    for f, m in zip(feats, g["feat_means"]):
        assert abs(f.mean().item() - m) < 1e-5


@pytest.mark.parametrize("norm", ["group", "batch", "none"])
def test_unet_norm_options(norm):
    """model.backbone.norm = group / batch (eval, running statistics) / anything else (Identity) — unet.py:29-41."""
    g = load("unet_norm_" + norm)
    y = OM.unet3d_forward(g["state_dict"], g["x"], norm=norm)
    _close(y, g["logits"], atol=5e-5, rtol=1e-4)
    if norm != "none":   # the kind is also recognisable from the state_dict alone
        _close(OM.unet3d_forward(g["state_dict"], g["x"]), g["logits"], atol=5e-5, rtol=1e-4)


def test_suv_guided_attention():
    g = load("suv_guided_attention")
    _close(OM.suv_guided_attention(g["state_dict"], g["ct"], g["pet"]), g["y"], atol=5e-5, rtol=1e-4)
    _close(OM.suv_guided_attention(g["state_dict"], g["ct"], g["pet_same"]), g["y_same"], atol=5e-5, rtol=1e-4)


def test_convblock_gelu_group_option():
    g = load("convblock_gelu_group")
    sd = {"b." + k: v for k, v in g["state_dict"].items()}
    _close(OM.conv_block3d(sd, "b", g["x"], activation="gelu", norm="group"), g["y"], atol=5e-5, rtol=1e-4)


def test_convblock_leaky_relu_option():
    g = load("convblock_leaky")
    sd = {"b." + k: v for k, v in g["state_dict"].items()}
    _close(OM.conv_block3d(sd, "b", g["x"], activation="leaky_relu"), g["y"])


@pytest.mark.parametrize("name,fusion", [("dual_attention_2", "attention"), ("dual_concat_2", "concat"),
                                         ("dual_cross_attention_2", "cross_attention"), ("dual_add_2", "add"),
                                         ("dual_attention_4", "attention")])
def test_dual_encoder_fusions(name, fusion):
    g = load(name)
    assert g["config"]["model"]["fusion"]["type"] == fusion
    _close(OM.dual_encoder_forward(g["state_dict"], g["x"], fusion), g["logits"])


def test_cross_attention_fusion_modules():
    g = load("cross_attention_fusion")
    _close(OM.cross_attention_fusion(g["state_dict"], g["q"], g["kv"], g["num_heads"]), g["y"], atol=5e-5)
    g = load("bidirectional_cross_attention")
    _close(OM.bidirectional_cross_attention(g["state_dict"], g["f1"], g["f2"]), g["y"], atol=5e-5)
    g = load("attention_fusion")
    _close(OM.attention_fusion(g["state_dict"], g["feats"]), g["y"])


def test_loss_known_answers():
    g = load("loss_kat_seed7")
    # SURVEY.md §4 known answers, measured on the reference by the survey
    assert abs(g["dicece"] - 1.0275284) < 1e-6 and abs(g["dice"] - 0.5845465) < 1e-6 and abs(g["ce"] - 1.4705102) < 1e-6
    tot, d, c = OL.dice_ce_loss(g["logits"], g["target"])
    assert abs(tot.item() - g["dicece"]) < 1e-6 and abs(d.item() - g["dice"]) < 1e-6 and abs(c.item() - g["ce"]) < 1e-6


def test_losses_and_gradients():
    g = load("losses")
    lg, tg, R = g["logits"], g["target"], g["results"]
    assert abs(OL.dice_ce_loss(lg, tg)[0].item() - R["dicece"]["value"]) < 1e-6
    assert abs(OL.dice_ce_loss(lg, tg)[0].item() - R["get_loss_default"]) < 1e-6
    assert abs(OL.dice_loss(lg, tg).item() - R["dice"]["value"]) < 1e-6
    assert abs(OL.dice_loss(lg, tg, include_background=False).item() - R["dice_nobg"]["value"]) < 1e-6
    assert abs(OL.focal_loss(lg, tg).item() - R["focal"]["value"]) < 1e-6
    assert abs(OL.tversky_loss(lg, tg, 0.3, 0.7).item() - R["tversky"]["value"]) < 1e-6
    w = torch.tensor([0.5, 1, 1, 2, 1, 1, 3, 1.0])
    assert abs(OL.dice_ce_loss(lg, tg, 0.3, 0.7, class_weights=w)[0].item() - R["dicece_w"]["value"]) < 1e-6
    gr = OL.dice_ce_grad(lg, tg).float()
    _close(gr, R["dicece"]["grad"], atol=1e-9, rtol=1e-4)


def test_dice_metric():
    g = load("dice_metric")
    # accumulate over both updates like DiceMetric does: concatenate along batch
    p = torch.cat(g["pred"])
    t = torch.cat(g["target"])
    r = OL.dice_metric(p, t, 8)
    assert abs(r["dice"] - g["dice"]) < 1e-6
    assert all(abs(a - b) < 1e-6 for a, b in zip(r["dice_per_class"], g["dice_per_class"]))
