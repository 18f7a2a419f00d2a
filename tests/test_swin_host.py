"""CPU tests of the SwinUNETR row (a19 / N2): the drop-in parameter tree, the factory wiring and the internal consistency
of the oracle restatement (PARITY UNPINNED: MONAI is absent; see oracle/swin_unetr.py)."""
import pytest
import torch

import mmseg_b200  # noqa: F401
from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR, build_swin_unetr
from oracle import swin_unetr as O


def test_state_dict_has_monai_names_and_size():
    m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48)
    sd = m.state_dict()
    for key, shape in {
        "model.swinViT.patch_embed.proj.weight": (48, 2, 2, 2, 2),
        "model.swinViT.layers1.0.blocks.1.attn.relative_position_bias_table": (2197, 3),
        "model.swinViT.layers1.0.blocks.0.attn.relative_position_index": (343, 343),
        "model.swinViT.layers3.0.blocks.0.attn.qkv.weight": (576, 192),
        "model.swinViT.layers2.0.blocks.1.mlp.linear1.weight": (384, 96),
        "model.swinViT.layers4.0.downsample.reduction.weight": (768, 3072),
        "model.swinViT.layers4.0.downsample.norm.bias": (3072,),
        "model.encoder1.layer.conv1.conv.weight": (48, 2, 3, 3, 3),
        "model.encoder1.layer.conv3.conv.weight": (48, 2, 1, 1, 1),
        "model.encoder10.layer.conv2.conv.weight": (768, 768, 3, 3, 3),
        "model.decoder5.transp_conv.conv.weight": (768, 384, 2, 2, 2),
        "model.decoder5.conv_block.conv3.conv.weight": (384, 768, 1, 1, 1),
        "model.decoder1.conv_block.conv1.conv.weight": (48, 96, 3, 3, 3),
        "model.out.conv.conv.bias": (8,),
    }.items():
        assert tuple(sd[key].shape) == shape, key
    assert "model.encoder2.layer.conv3.conv.weight" not in sd          # no residual conv without a channel change
    n_params = sum(p.numel() for p in m.parameters())
    assert abs(n_params / 1e6 - 62.19) < 0.01                           # the published size of SwinUNETR-48
    assert m.encoder_channels == [48, 96, 192, 384, 768]
    assert torch.equal(sd["model.swinViT.layers1.0.blocks.0.attn.relative_position_index"], O.relative_position_index((7, 7, 7)))


def test_factory_builds_swin_unetr_and_module_fails_loudly_on_cpu():
    from mmseg_b200.src.models.build import build_model, MODEL_REGISTRY
    assert MODEL_REGISTRY["swin_unetr"] is build_swin_unetr
    cfg = {"model": {"name": "swin_unetr", "in_channels": 1, "out_channels": 8, "backbone": {"feature_size": 48}},
           "data": {"modalities": ["ct", "pet"]}, "hardware": {"device": "cpu"}}
    model = build_model(cfg)
    assert cfg["model"]["in_channels"] == 2 and model.backbone.in_channels == 2   # reference build.py:98-99
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 2, 32, 32, 32))
    with pytest.raises(NotImplementedError):
        SwinUNETR(use_v2=True)
    with pytest.raises(NotImplementedError):
        SwinUNETR(feature_size=24, num_heads=(3, 6, 12, 24))


def test_oracle_window_partition_roundtrip_and_mask():
    x = torch.randn(2, 14, 7, 21, 5)
    ws = (7, 7, 7)
    w = O.window_partition(x, ws)
    assert w.shape == (2 * 2 * 1 * 3, 343, 5)
    assert torch.equal(O.window_reverse(w.view(-1, 7, 7, 7, 5), ws, (2, 14, 7, 21)), x)
    mask = O.compute_mask((14, 14, 14), ws, (3, 3, 3))
    assert mask.shape == (8, 343, 343) and set(mask.unique().tolist()) == {-100.0, 0.0}
    assert (mask[0] == 0).all()                       # the first window lies inside one region
    assert (mask.diagonal(dim1=1, dim2=2) == 0).all() and torch.equal(mask, mask.transpose(1, 2))


def test_oracle_shifted_window_attention_against_brute_force():
    """The oracle's pad -> roll -> partition -> (bias, mask) -> reverse -> roll pipeline equals a direct evaluation over
    all token pairs of the padded grid: tokens attend iff they share a shifted window and a shift region, with the bias
    taken from their 3-D offset."""
    torch.manual_seed(0)
    dims, heads, window = (8, 6, 9), 2, (7, 7, 7)
    C_ = heads * 16
    sd = {"a.qkv.weight": torch.randn(3 * C_, C_) * 0.2, "a.qkv.bias": torch.randn(3 * C_) * 0.2,
          "a.proj.weight": torch.eye(C_), "a.proj.bias": torch.zeros(C_),
          "a.relative_position_bias_table": torch.randn(13 ** 3, heads)}
    x = torch.randn(1, *dims, C_, dtype=torch.float64)
    shift = (3, 3, 3)
    ws, ss = O.get_window_size(dims, window, shift)
    assert ws == (7, 6, 7) and ss == (3, 0, 3)
    pads = [(ws[i] - dims[i] % ws[i]) % ws[i] for i in range(3)]
    xp = torch.nn.functional.pad(x, (0, 0, 0, pads[2], 0, pads[1], 0, pads[0]))
    P = xp.shape[1:4]
    xr = torch.roll(xp, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    mask = O.compute_mask(P, ws, ss).double()
    index = O.relative_position_index(window)
    aw = O.window_attention(sd, "a", O.window_partition(xr, ws), heads, mask, index)
    got = torch.roll(O.window_reverse(aw.view(-1, *ws, C_), ws, (1, *P)), shifts=ss, dims=(1, 2, 3))[0]
    # brute force on the rolled grid
    qkv = (xr[0].reshape(-1, C_) @ sd["a.qkv.weight"].double().t() + sd["a.qkv.bias"].double()).view(-1, 3, heads, 16)
    g = torch.stack(torch.meshgrid(*[torch.arange(p) for p in P], indexing="ij"), -1).reshape(-1, 3)
    wid = torch.stack([g[:, i] // ws[i] for i in range(3)], -1)
    loc = torch.stack([g[:, i] % ws[i] for i in range(3)], -1)
    flat = (loc[:, 0] * ws[1] + loc[:, 1]) * ws[2] + loc[:, 2]         # token index inside its window
    reg = torch.zeros(len(g), dtype=torch.long)
    for i in range(3):
        r = torch.zeros(len(g), dtype=torch.long)
        if ss[i] > 0:
            r = (g[:, i] >= P[i] - ws[i]).long() + (g[:, i] >= P[i] - ss[i]).long()
        reg = reg * 3 + r
    same_win = (wid[:, None, :] == wid[None, :, :]).all(-1)
    same_reg = reg[:, None] == reg[None, :]
    bias_idx = index[flat][:, flat]                                     # MONAI: index[:n, :n] by in-window token index
    tab = sd["a.relative_position_bias_table"].double()
    out = torch.zeros(len(g), C_, dtype=torch.float64)
    for h in range(heads):
        q, k, v = qkv[:, 0, h] * 0.25, qkv[:, 1, h], qkv[:, 2, h]
        s = q @ k.t() + tab[bias_idx.reshape(-1), h].view(len(g), len(g))
        s = s + torch.where(same_reg, 0.0, -100.0)
        s = s.masked_fill(~same_win, float("-inf"))
        out[:, h * 16:(h + 1) * 16] = s.softmax(-1) @ v
    want = torch.roll(out.view(*P, C_).unsqueeze(0), shifts=ss, dims=(1, 2, 3))[0]
    assert (got - want).abs().max().item() < 1e-10


def test_oracle_forward_shapes_and_determinism():
    torch.manual_seed(0)
    m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).eval()
    sd = m.state_dict()
    x = torch.randn(1, 2, 64, 64, 64)
    y, hs = O.swin_unetr_forward(sd, x, return_hidden=True)
    assert y.shape == (1, 8, 64, 64, 64) and torch.isfinite(y).all()
    assert [tuple(h.shape[1:]) for h in hs] == [(48, 32, 32, 32), (96, 16, 16, 16), (192, 8, 8, 8), (384, 4, 4, 4), (768, 2, 2, 2)]
    for h in hs[:4]:   # proj_out: LayerNorm over channels without affine
        assert h.mean(1).abs().max().item() < 1e-4 and (h.var(1, unbiased=False) - 1).abs().max().item() < 1e-2
    assert torch.equal(y, O.swin_unetr_forward(sd, x))
