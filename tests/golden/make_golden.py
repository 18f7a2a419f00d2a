"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
It imports the reference's modules read-only (src.models.build, src.models.fusion.attention_fusion,
src.trainer.losses, src.trainer.metrics), runs them on CPU fp32 with fixed seeds and stores inputs, parameters and
outputs as small .pt files.  The reference has no tests or golden vectors of its own, so these
self-generated vectors are what pins the oracle (oracle/) and the CUDA path on the GPU box, where /root/reference
does not exist.
"""
import copy
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _cfg(name, fusion, modalities, features, out_channels=8):
    return {
        "model": {"name": name, "in_channels": len(modalities), "out_channels": out_channels,
                  "backbone": {"features": list(features), "norm": "instance"},
                  "fusion": {"type": fusion}, "head": {"dropout": 0.0}},
        "data": {"modalities": list(modalities)},
        "hardware": {"device": "cpu"},
    }


def main():
    sys.path.insert(0, REF)
    from src.models.build import build_model
    from src.models.fusion.attention_fusion import AttentionFusion, CrossAttentionFusion, BidirectionalCrossAttention
    from src.trainer.losses import get_loss, DiceLoss, DiceCELoss, FocalLoss, TverskyLoss
    from src.trainer.metrics import DiceMetric

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}

    # ---- UNet3D (early fusion, CT+PET), small features so the fixture stays small
    torch.manual_seed(0)
    cfg = _cfg("unet", "early", ["CT", "PET"], [16, 32, 64])
    m = build_model(copy.deepcopy(cfg)).eval()
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 2, 32, 32, 32, generator=g)
    with torch.no_grad():
        y, feats = m(x, return_features=True)
    out["unet_small"] = {"config": cfg, "state_dict": m.state_dict(), "x": x, "logits": y,
                         "feat_means": [f.mean().item() for f in feats]}

    # ---- UNet3D leaky_relu variant is not reachable from the config; ConvBlock3D option checked directly
    from src.models.backbones.unet import ConvBlock3D
    torch.manual_seed(1)
    blk = ConvBlock3D(16, 32, activation="leaky_relu").eval()
    xb = torch.randn(2, 16, 8, 12, 16, generator=g)
    with torch.no_grad():
        yb = blk(xb)
    out["convblock_leaky"] = {"state_dict": blk.state_dict(), "x": xb, "y": yb}

    # ---- DualEncoder, every fusion branch of _fuse_features (dual_encoder.py:167-199)
    for fusion, mods in (("attention", ["CT", "PET"]), ("concat", ["CT", "PET"]), ("cross_attention", ["CT", "PET"]),
                         ("add", ["CT", "PET"]), ("attention", ["CT", "PET", "MRI", "US"])):
        torch.manual_seed(2)
        cfg = _cfg("dual_encoder", fusion, mods, [16, 32])
        m = build_model(copy.deepcopy(cfg)).eval()
        x = torch.randn(2 if len(mods) == 2 else 1, len(mods), 16, 16, 16, generator=g)
        with torch.no_grad():
            y = m(x)
        out[f"dual_{fusion}_{len(mods)}"] = {"config": cfg, "state_dict": m.state_dict(), "x": x, "logits": y}

    # ---- fusion modules (attention_fusion.py)
    torch.manual_seed(3)
    caf = CrossAttentionFusion(32, num_heads=4).eval()
    q, kv = torch.randn(2, 32, 4, 6, 5, generator=g), torch.randn(2, 32, 4, 6, 5, generator=g)
    with torch.no_grad():
        yc = caf(q, kv)
    out["cross_attention_fusion"] = {"state_dict": caf.state_dict(), "q": q, "kv": kv, "y": yc, "num_heads": 4}
    torch.manual_seed(4)
    bca = BidirectionalCrossAttention(32, num_heads=4).eval()
    with torch.no_grad():
        yb2 = bca(q, kv)
    out["bidirectional_cross_attention"] = {"state_dict": bca.state_dict(), "f1": q, "f2": kv, "y": yb2}
    torch.manual_seed(5)
    af = AttentionFusion(16, 2).eval()
    f1, f2 = torch.randn(2, 16, 4, 4, 4, generator=g), torch.randn(2, 16, 4, 4, 4, generator=g)
    with torch.no_grad():
        ya = af([f1, f2])
    out["attention_fusion"] = {"state_dict": af.state_dict(), "feats": [f1, f2], "y": ya}

    # ---- losses (losses.py) with gradients
    torch.manual_seed(7)
    lg = torch.randn(2, 3, 2, 2, 2)
    tg = torch.randint(0, 3, (2, 2, 2, 2))
    kat = {"logits": lg, "target": tg, "dicece": DiceCELoss()(lg, tg).item(), "dice": DiceLoss()(lg, tg).item(),
           "ce": torch.nn.functional.cross_entropy(lg, tg).item()}
    out["loss_kat_seed7"] = kat  # 1.0275284 / 0.5845465 / 1.4705102
    lg = torch.randn(2, 8, 12, 10, 14, generator=g)
    tg = torch.randint(0, 8, (2, 12, 10, 14), generator=g)
    losses = {}
    for name, fn in (("dicece", DiceCELoss()), ("dice", DiceLoss()), ("dice_nobg", DiceLoss(include_background=False)),
                     ("focal", FocalLoss()), ("tversky", TverskyLoss(alpha=0.3, beta=0.7)),
                     ("dicece_w", DiceCELoss(dice_weight=0.3, ce_weight=0.7,
                                             class_weights=torch.tensor([0.5, 1, 1, 2, 1, 1, 3, 1.0])))):
        z = lg.clone().requires_grad_(True)
        v = fn(z, tg)
        v.backward()
        losses[name] = {"value": v.item(), "grad": z.grad.clone()}
    lcfg = {"training": {"loss": {"name": "dice_ce", "dice_weight": 0.5, "ce_weight": 0.5}}}
    losses["get_loss_default"] = get_loss(lcfg)(lg, tg).item()
    out["losses"] = {"logits": lg, "target": tg, "results": losses}

    # ---- DiceMetric (metrics.py:11-88) accumulated over two updates
    dm = DiceMetric(num_classes=8)
    p1, t1 = torch.randint(0, 8, (2, 6, 6, 6), generator=g), torch.randint(0, 8, (2, 6, 6, 6), generator=g)
    p2, t2 = torch.randint(0, 8, (1, 6, 6, 6), generator=g), torch.randint(0, 8, (1, 6, 6, 6), generator=g)
    dm.update(p1, t1)
    dm.update(p2, t2)
    r = dm.compute()
    out["dice_metric"] = {"pred": [p1, p2], "target": [t1, t2],
                          "dice": float(r["dice"]), "dice_per_class": [float(v) for v in r["dice_per_class"]]}

    # ---- DeepSupervisionHead (heads/segmentation.py:62-115): 3 scales, trilinear align_corners=True resize to the finest
    from src.models.heads.segmentation import DeepSupervisionHead
    torch.manual_seed(21)
    ds = DeepSupervisionHead([16, 32, 64], 5).eval()
    g2 = torch.Generator().manual_seed(22)
    feats = [torch.randn(2, 16, 12, 10, 14, generator=g2), torch.randn(2, 32, 6, 5, 7, generator=g2),
             torch.randn(2, 64, 3, 3, 4, generator=g2)]
    with torch.no_grad():
        outs = ds(feats, target_size=(12, 10, 14))
    out["deep_supervision"] = {"state_dict": ds.state_dict(), "features": feats, "target_size": (12, 10, 14),
                               "outputs": [o.clone() for o in outs]}

    # ---- UNet3D with model.backbone.norm = group / batch / none (unet.py:29-41,224): non-trivial affine parameters and,
    # for batch, running statistics populated by two train-mode forwards before the eval-mode reference output
    for norm in ("group", "batch", "none"):
        torch.manual_seed(30)
        cfgn = _cfg("unet", "early", ["CT", "PET"], [16, 32])
        cfgn["model"]["backbone"]["norm"] = norm
        mn = build_model(copy.deepcopy(cfgn))
        gn = torch.Generator().manual_seed(31)
        with torch.no_grad():
            for prm_name, prm in mn.named_parameters():
                if ".norm" in prm_name:      # gamma around 1, beta around 0 (defaults are exactly 1 / 0)
                    prm.add_(0.3 * torch.randn(prm.shape, generator=gn))
            if norm == "batch":
                mn.train()
                for _ in range(2):
                    mn(torch.randn(2, 2, 16, 16, 16, generator=gn) * 1.5 + 0.3)
        mn.eval()
        xn = torch.randn(2, 2, 16, 16, 16, generator=gn)
        with torch.no_grad():
            yn = mn(xn)
        out["unet_norm_" + norm] = {"config": cfgn, "state_dict": mn.state_dict(), "x": xn, "logits": yn}

    # ---- ConvBlock3D(activation="gelu") with group norm (unet.py:36-38,47-48): options only reachable by direct construction
    torch.manual_seed(40)
    blkg = ConvBlock3D(16, 32, norm="group", activation="gelu").eval()
    gg = torch.Generator().manual_seed(41)
    xg = torch.randn(2, 16, 8, 10, 12, generator=gg)
    with torch.no_grad():
        yg = blkg(xg)
    out["convblock_gelu_group"] = {"state_dict": blkg.state_dict(), "x": xg, "y": yg}

    # ---- SUVGuidedAttention (attention_fusion.py:219-295), PET given at half resolution (exercises the trilinear resize)
    from src.models.fusion.attention_fusion import SUVGuidedAttention
    torch.manual_seed(50)
    suv = SUVGuidedAttention(32, suv_threshold=1.2).eval()
    gs = torch.Generator().manual_seed(51)
    ctf = torch.randn(2, 32, 8, 10, 12, generator=gs)
    pets = torch.rand(2, 1, 4, 5, 6, generator=gs) * 4.0
    with torch.no_grad():
        ys = suv(ctf, pets)
        ys_same = suv(ctf, torch.rand(2, 1, 8, 10, 12, generator=torch.Generator().manual_seed(52)) * 4.0)
    out["suv_guided_attention"] = {"state_dict": suv.state_dict(), "ct": ctf, "pet": pets, "y": ys,
                                   "pet_same": torch.rand(2, 1, 8, 10, 12, generator=torch.Generator().manual_seed(52)) * 4.0,
                                   "y_same": ys_same}

    only = set(sys.argv[1:])      # python make_golden.py [name ...]: write only these fixtures (all are seeded)
    for k, v in out.items():
        if only and k not in only:
            continue
        torch.save(v, os.path.join(HERE, k + ".pt"))
        print(k, f"{os.path.getsize(os.path.join(HERE, k + '.pt')) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
