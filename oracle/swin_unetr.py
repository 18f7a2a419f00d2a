"""Oracle (CPU, torch.nn.functional) restatement of SwinUNETR — BASELINE.json configs[3], SURVEY.md row a19 / N2.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

*** PARITY UNPINNED ***  The reference's `SwinUNETR` (src/models/backbones/swin_unetr.py:20-200) is a thin wrapper:
its ctor builds `monai.networks.nets.SwinUNETR` (swin_unetr.py:80-96) and `forward` is `self.model(x)` (:117).  The
arithmetic therefore lives in MONAI (`requirements.txt:7`: `monai>=1.3.0`, a floor not a pin), which is neither vendored
under /root/reference nor installed in this image, and the reference has no test or golden vector for this path.  What
follows restates the published MONAI 1.3 algorithm (`monai/networks/nets/swin_unetr.py`, `monai/networks/blocks/
{unetr_block,dynunet_block,mlp,patchembedding}.py`) from its specification:

  SwinUNETR.forward      hidden = swinViT(x, normalize);  enc0 = encoder1(x);  enc1..3 = encoder2..4(hidden[0..2]);
                         dec4 = encoder10(hidden[4]);  dec3 = decoder5(dec4, hidden[3]);  dec2 = decoder4(dec3, enc3);
                         dec1 = decoder3(dec2, enc2);  dec0 = decoder2(dec1, enc1);  out = decoder1(dec0, enc0);
                         logits = out(out)                       (1x1x1 conv with bias)
  SwinTransformer        patch_embed = Conv3d(in, F, k2, s2, bias);  4 stages of BasicLayer(dim = F * 2^i): `depth`
                         SwinTransformerBlocks (shift 0 / window//2 alternating) then PatchMerging; every hidden state is
                         returned through proj_out = LayerNorm over channels WITHOUT affine (F.layer_norm(x, [C]))
  SwinTransformerBlock   x = x + attn(window_partition(roll(pad(norm1(x))))) ;  x = x + mlp(norm2(x))  (mlp: Linear 4x,
                         GELU, Linear);  zero padding AFTER norm1 up to a multiple of the window, cyclic shift on the
                         padded grid, attention mask only in shifted blocks; when a spatial size <= window the window
                         shrinks to it and the shift becomes 0 (get_window_size)
  WindowAttention        qkv Linear (bias), q * head_dim^-0.5, + relative_position_bias_table[index[:n, :n]], + mask
                         (0 / -100), softmax, @ v, proj Linear.  index[:n, :n] is sliced from the 7^3 index also when the
                         window shrank (MONAI does exactly this)
  PatchMerging (v1)      the legacy 3-D gather order x0..x7 = [0,0,0],[1,0,0],[0,1,0],[0,0,1],[1,0,1],[0,1,0],[0,0,1],
                         [1,1,1] (offsets along (d, h, w); two octants appear twice, two never — kept by MONAI for
                         checkpoint compatibility), LayerNorm(8C) with affine, Linear(8C -> 2C, no bias)
  UnetrBasicBlock /      UnetResBlock: conv3(k3, no bias) - InstanceNorm3d(affine=False) - LeakyReLU(0.01) - conv3 - IN,
  UnetrUpBlock           residual (1x1x1 conv + IN when channels change), LeakyReLU(0.01);  UpBlock = ConvTranspose3d(k2,
                         s2, no bias), cat([up, skip]), UnetResBlock

Parameter names are MONAI's (`swinViT.layers1.0.blocks.0.attn.qkv.weight`, `encoder1.layer.conv1.conv.weight`,
`decoder5.transp_conv.conv.weight`, `out.conv.conv.weight` ...), under the reference wrapper's `model.` attribute.
Dropout / drop-path are the identity (eval, and the reference passes rate 0 by default).
"""
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _p(sd: SD, key: str, dtype) -> Tensor:
    return sd[key].detach().to("cpu", dtype)


def relative_position_index(ws: Sequence[int]) -> Tensor:
    """WindowAttention.__init__: pairwise relative offsets of the ws[0]*ws[1]*ws[2] window tokens -> table row."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]), indexing="ij"))
    cf = coords.flatten(1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 2] += ws[2] - 1
    rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
    rel[:, :, 1] *= 2 * ws[2] - 1
    return rel.sum(-1)


def get_window_size(x_size, window_size, shift_size):
    ws, ss = list(window_size), list(shift_size)
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            ws[i] = x_size[i]
            ss[i] = 0
    return tuple(ws), tuple(ss)


def window_partition(x: Tensor, ws) -> Tensor:
    b, d, h, w, c = x.shape
    x = x.view(b, d // ws[0], ws[0], h // ws[1], ws[1], w // ws[2], ws[2], c)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, ws[0] * ws[1] * ws[2], c)


def window_reverse(windows: Tensor, ws, dims) -> Tensor:
    b, d, h, w = dims
    x = windows.view(b, d // ws[0], h // ws[1], w // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(b, d, h, w, -1)


def compute_mask(dims, ws, ss) -> Tensor:
    d, h, w = dims
    img = torch.zeros((1, d, h, w, 1))
    cnt = 0
    for ds in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for hs in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for wsl in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, ds, hs, wsl, :] = cnt
                cnt += 1
    mw = window_partition(img, ws).squeeze(-1)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def window_attention(sd: SD, prefix: str, x: Tensor, heads: int, mask, index: Tensor) -> Tensor:
    dt = x.dtype
    b, n, c = x.shape
    hd = c // heads
    qkv = F.linear(x, _p(sd, prefix + ".qkv.weight", dt), _p(sd, prefix + ".qkv.bias", dt) if prefix + ".qkv.bias" in sd else None)
    qkv = qkv.reshape(b, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    table = _p(sd, prefix + ".relative_position_bias_table", dt)
    bias = table[index[:n, :n].reshape(-1)].reshape(n, n, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(b // nw, nw, heads, n, n) + mask.to(dt).unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = attn.softmax(-1)
    x = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(x, _p(sd, prefix + ".proj.weight", dt), _p(sd, prefix + ".proj.bias", dt))


def swin_block(sd: SD, prefix: str, x: Tensor, heads: int, window, shift, mask_matrix, index: Tensor) -> Tensor:
    """SwinTransformerBlock.forward on x [b, d, h, w, c]."""
    dt = x.dtype
    b, d, h, w, c = x.shape
    shortcut = x
    x = F.layer_norm(x, [c], _p(sd, prefix + ".norm1.weight", dt), _p(sd, prefix + ".norm1.bias", dt), 1e-5)
    ws, ss = get_window_size((d, h, w), window, shift)
    pd, ph, pw = (ws[0] - d % ws[0]) % ws[0], (ws[1] - h % ws[1]) % ws[1], (ws[2] - w % ws[2]) % ws[2]
    x = F.pad(x, (0, 0, 0, pw, 0, ph, 0, pd))
    _, dp, hp, wp, _ = x.shape
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        mask = mask_matrix
    else:
        mask = None
    aw = window_attention(sd, prefix + ".attn", window_partition(x, ws), heads, mask, index)
    x = window_reverse(aw.view(-1, ws[0], ws[1], ws[2], c), ws, (b, dp, hp, wp))
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=ss, dims=(1, 2, 3))
    x = shortcut + x[:, :d, :h, :w, :]
    y = F.layer_norm(x, [c], _p(sd, prefix + ".norm2.weight", dt), _p(sd, prefix + ".norm2.bias", dt), 1e-5)
    y = F.linear(y, _p(sd, prefix + ".mlp.linear1.weight", dt), _p(sd, prefix + ".mlp.linear1.bias", dt))
    y = F.linear(F.gelu(y), _p(sd, prefix + ".mlp.linear2.weight", dt), _p(sd, prefix + ".mlp.linear2.bias", dt))
    return x + y


MERGE_OFFSETS = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1))


def patch_merging(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """PatchMerging.forward (the v1 "merging" layer) on [b, d, h, w, c] with even d, h, w."""
    dt = x.dtype
    b, d, h, w, c = x.shape
    x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2, 0, d % 2))
    x = torch.cat([x[:, i::2, j::2, k::2, :] for (i, j, k) in MERGE_OFFSETS], -1)
    x = F.layer_norm(x, [8 * c], _p(sd, prefix + ".norm.weight", dt), _p(sd, prefix + ".norm.bias", dt), 1e-5)
    return F.linear(x, _p(sd, prefix + ".reduction.weight", dt))


def basic_layer(sd: SD, prefix: str, x: Tensor, depth: int, heads: int, window, index: Tensor) -> Tensor:
    """BasicLayer.forward: x [b, c, d, h, w] -> blocks -> PatchMerging -> [b, 2c, d/2, h/2, w/2]."""
    b, c, d, h, w = x.shape
    shift = tuple(i // 2 for i in window)
    ws, ss = get_window_size((d, h, w), window, shift)
    x = x.permute(0, 2, 3, 4, 1)
    dp, hp, wp = (-(-d // ws[0])) * ws[0], (-(-h // ws[1])) * ws[1], (-(-w // ws[2])) * ws[2]
    mask = compute_mask((dp, hp, wp), ws, ss)
    for i in range(depth):
        x = swin_block(sd, f"{prefix}.blocks.{i}", x, heads, window, (0, 0, 0) if i % 2 == 0 else shift, mask, index)
    x = patch_merging(sd, prefix + ".downsample", x)
    return x.permute(0, 4, 1, 2, 3)


def _proj_out(x: Tensor, normalize: bool) -> Tensor:
    if not normalize:
        return x
    c = x.shape[1]
    return F.layer_norm(x.permute(0, 2, 3, 4, 1), [c]).permute(0, 4, 1, 2, 3)


def swin_vit(sd: SD, prefix: str, x: Tensor, depths, heads, window=(7, 7, 7), normalize: bool = True) -> List[Tensor]:
    dt = x.dtype
    index = relative_position_index(window)
    x0 = F.conv3d(x, _p(sd, prefix + "patch_embed.proj.weight", dt), _p(sd, prefix + "patch_embed.proj.bias", dt), stride=2)
    outs = [_proj_out(x0, normalize)]
    cur = x0
    for i in range(4):
        cur = basic_layer(sd, f"{prefix}layers{i + 1}.0", cur.contiguous(), depths[i], heads[i], window, index)
        outs.append(_proj_out(cur, normalize))
    return outs


def _inorm(x: Tensor) -> Tensor:
    return F.instance_norm(x, eps=1e-5)


def unet_res_block(sd: SD, prefix: str, x: Tensor) -> Tensor:
    dt = x.dtype
    out = F.conv3d(x, _p(sd, prefix + ".conv1.conv.weight", dt), None, padding=1)
    out = F.leaky_relu(_inorm(out), 0.01)
    out = _inorm(F.conv3d(out, _p(sd, prefix + ".conv2.conv.weight", dt), None, padding=1))
    res = x
    if prefix + ".conv3.conv.weight" in sd:
        res = _inorm(F.conv3d(x, _p(sd, prefix + ".conv3.conv.weight", dt), None))
    return F.leaky_relu(out + res, 0.01)


def unetr_up_block(sd: SD, prefix: str, x: Tensor, skip: Tensor) -> Tensor:
    dt = x.dtype
    up = F.conv_transpose3d(x, _p(sd, prefix + ".transp_conv.conv.weight", dt), None, stride=2)
    return unet_res_block(sd, prefix + ".conv_block", torch.cat((up, skip), dim=1))


def swin_unetr_forward(sd: SD, x: Tensor, prefix: str = "model.", depths=(2, 2, 2, 2), heads=(3, 6, 12, 24),
                       normalize: bool = True, dtype=torch.float32, return_hidden: bool = False):
    """SwinUNETR.forward (reference wrapper swin_unetr.py:102-117 -> MONAI SwinUNETR.forward)."""
    x = x.detach().to("cpu", dtype)
    hs = swin_vit(sd, prefix + "swinViT.", x, depths, heads, (7, 7, 7), normalize)
    enc0 = unet_res_block(sd, prefix + "encoder1.layer", x)
    enc1 = unet_res_block(sd, prefix + "encoder2.layer", hs[0])
    enc2 = unet_res_block(sd, prefix + "encoder3.layer", hs[1])
    enc3 = unet_res_block(sd, prefix + "encoder4.layer", hs[2])
    dec4 = unet_res_block(sd, prefix + "encoder10.layer", hs[4])
    dec3 = unetr_up_block(sd, prefix + "decoder5", dec4, hs[3])
    dec2 = unetr_up_block(sd, prefix + "decoder4", dec3, enc3)
    dec1 = unetr_up_block(sd, prefix + "decoder3", dec2, enc2)
    dec0 = unetr_up_block(sd, prefix + "decoder2", dec1, enc1)
    out = unetr_up_block(sd, prefix + "decoder1", dec0, enc0)
    logits = F.conv3d(out, _p(sd, prefix + "out.conv.conv.weight", dtype), _p(sd, prefix + "out.conv.conv.bias", dtype))
    return (logits, hs) if return_hidden else logits
