"""Forward engine of SwinUNETR (BASELINE.json configs[3]; reference src/models/backbones/swin_unetr.py:80-117, which
delegates to monai.networks.nets.SwinUNETR — the launch sequence below follows MONAI 1.3's SwinUNETR.forward).

Everything is a C-ABI kernel launch on the current stream:
  * linear layers (qkv, proj, mlp, patch-merging reduction) and every conv of the UNETR encoder / decoder blocks run on
    the tcgen05 conv kernel (1x1x1 / 3x3x3 implicit GEMM, ConvTranspose as GEMM + pixel shuffle into the concat buffer);
  * shifted-window attention (padding, cyclic shift, window partition / reverse, relative-position bias, shift mask and
    softmax) is one kernel (swin.cu);
  * the token residual stream is blocked fp32; `swin_layernorm` fuses the residual add with the next LayerNorm;
  * UnetResBlock = conv -> IN statistics in the conv epilogue -> apply(+LeakyReLU 0.01); its tail
    LeakyReLU(IN(conv2) + IN(conv3) | x) is one kernel (`instnorm_residual_act`).
Numeric format: fp16 (default) or bf16 operands, fp32 accumulation, fp32 raw conv outputs and residual stream.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import kernels as K
from .engine import _Workspace, _param_version
from .kernels import Blocked, PackedConv
from .numerics import NumericMode

Tensor = torch.Tensor

SWIN_MODES = {
    "fp16": NumericMode("fp16", _lib.FMT_FP16, False, False, True),
    "bf16": NumericMode("bf16", _lib.FMT_BF16, False, False, True),
}
LRELU_SLOPE = 0.01   # MONAI UnetResBlock: act_name=("leakyrelu", {"negative_slope": 0.01})


def _lin_w(lin) -> Tensor:
    return lin.weight.detach().reshape(lin.out_features, lin.in_features, 1, 1, 1)


class SwinUNETREngine:
    def __init__(self, net, mode: str = "fp16", weights_from: Optional["SwinUNETREngine"] = None):
        """weights_from: another engine of the same net and mode whose packed weights this one reads (the extra batch
        slots of the sliding-window inferer)."""
        if mode not in SWIN_MODES:
            raise ValueError(f"SwinUNETR numeric mode {mode!r}; choose from {sorted(SWIN_MODES)}")
        self.net = self.module = net       # `module`: the name the sliding-window inferer uses
        self.mode = mode
        self._weights_from = weights_from
        self.nm = SWIN_MODES[mode]
        self._packed: Optional[Dict[str, PackedConv]] = None
        self._packed_version = None
        self._bufs: Dict[Tuple, Dict[str, object]] = {}
        self._ws: Optional[_Workspace] = None
        self._ident: Dict[Tuple, Tensor] = {}
        self.launches = 0

    # ---------------------------------------------------------------- weights
    def _pack(self) -> Dict[str, PackedConv]:
        if self._weights_from is not None:
            return self._weights_from._pack()
        net = self.net
        ver = _param_version(list(net.parameters()))
        if self._packed is not None and ver == self._packed_version:
            return self._packed
        nm = self.nm
        P: Dict[str, PackedConv] = {}
        vit = net.swinViT
        for s in range(4):
            layer = getattr(vit, f"layers{s + 1}")[0]
            for j, blk in enumerate(layer.blocks):
                p = f"l{s}.b{j}."
                P[p + "qkv"] = K.pack_conv_weight(_lin_w(blk.attn.qkv), blk.attn.qkv.bias, nm)
                P[p + "proj"] = K.pack_conv_weight(_lin_w(blk.attn.proj), blk.attn.proj.bias, nm)
                P[p + "fc1"] = K.pack_conv_weight(_lin_w(blk.mlp.linear1), blk.mlp.linear1.bias, nm)
                P[p + "fc2"] = K.pack_conv_weight(_lin_w(blk.mlp.linear2), blk.mlp.linear2.bias, nm)
            P[f"l{s}.red"] = K.pack_conv_weight(_lin_w(layer.downsample.reduction), None, nm)

        def res(name: str, blk, segs):
            P[name + ".c1"] = K.pack_conv_weight(blk.conv1.conv.weight, None, nm, segs)
            P[name + ".c2"] = K.pack_conv_weight(blk.conv2.conv.weight, None, nm)
            if getattr(blk, "conv3", None) is not None:
                P[name + ".c3"] = K.pack_conv_weight(blk.conv3.conv.weight, None, nm, segs)

        F = net.feature_size
        res("enc1", net.encoder1.layer, [net.in_channels])
        res("enc2", net.encoder2.layer, [F])
        res("enc3", net.encoder3.layer, [2 * F])
        res("enc4", net.encoder4.layer, [4 * F])
        res("enc10", net.encoder10.layer, [16 * F])
        for name, dec, c in (("dec5", net.decoder5, 8 * F), ("dec4", net.decoder4, 4 * F), ("dec3", net.decoder3, 2 * F),
                             ("dec2", net.decoder2, F), ("dec1", net.decoder1, F)):
            P[name + ".up"] = K.pack_conv_weight(dec.transp_conv.conv.weight, None, nm, transposed=True)
            res(name, dec.conv_block, [c, c])
        self._packed, self._packed_version = P, ver
        return P

    # ---------------------------------------------------------------- buffers
    def _buffers(self, n: int, Z: int, Y: int, X: int, device) -> Dict[str, object]:
        key = (n, Z, Y, X)
        b = self._bufs.get(key)
        if b is not None:
            return b
        if any(d % 32 for d in (Z, Y, X)):
            raise NotImplementedError(f"SwinUNETR needs spatial sizes divisible by 32, got {(Z, Y, X)} (MONAI raises as well)")
        net, nm = self.net, self.nm
        F = net.feature_size
        f32 = lambda c, d: torch.empty((n, c // 8, d[0], d[1], d[2], 8), dtype=torch.float32, device=device)
        dims = [(Z >> l, Y >> l, X >> l) for l in range(6)]
        b = {"in": Blocked(n, (net.in_channels + 15) // 16 * 16, Z, Y, X, nm, device)}
        b["in"].t.zero_()
        b["x32"] = torch.empty((n, net.in_channels, Z, Y, X), dtype=torch.float32, device=device)   # sliding-window batches
        # token path: stage s works on C = F * 2^s channels at dims[s + 1]
        for s in range(4):
            C_, d = F << s, dims[s + 1]
            b[f"xs{s}"] = f32(C_, d)
            b[f"ln{s}"] = Blocked(n, C_, *d, nm, device)
            b[f"qkv{s}"] = Blocked(n, 3 * C_, *d, nm, device)
            b[f"att{s}"] = Blocked(n, C_, *d, nm, device)
            b[f"y{s}"] = f32(C_, d)
            b[f"h32_{s}"] = f32(4 * C_, d)
            b[f"h{s}"] = Blocked(n, 4 * C_, *d, nm, device)
            b[f"mg{s}"] = Blocked(n, 8 * C_, *dims[s + 2], nm, device)
        b["xs4"] = f32(16 * F, dims[5])
        # hidden states (LayerNorm'ed, no affine): 0..2 and 4 feed encoder blocks, 3 is decoder5's skip
        for i in (0, 1, 2, 4):
            b[f"hid{i}"] = Blocked(n, F << i, *dims[i + 1], nm, device)
        # concat buffers [up | skip] of the five decoder stages, at dims[4] ... dims[0]
        b["cat5"] = Blocked(n, 16 * F, *dims[4], nm, device)
        b["cat4"] = Blocked(n, 8 * F, *dims[3], nm, device)
        b["cat3"] = Blocked(n, 4 * F, *dims[2], nm, device)
        b["cat2"] = Blocked(n, 2 * F, *dims[1], nm, device)
        b["cat1"] = Blocked(n, 2 * F, *dims[0], nm, device)
        b["e10"] = Blocked(n, 16 * F, *dims[5], nm, device)
        for name, c, l in (("d5", 8 * F, 4), ("d4", 4 * F, 3), ("d3", 2 * F, 2), ("d2", F, 1), ("d1", F, 0)):
            b[name] = Blocked(n, c, *dims[l], nm, device)
            b["mid_" + name] = Blocked(n, c, *dims[l], nm, device)
        b["mid_e10"] = Blocked(n, 16 * F, *dims[5], nm, device)
        self._bufs[key] = b
        return b

    def input_buffer(self, n: int, Z: int, Y: int, X: int, device) -> Blocked:
        return self._buffers(n, Z, Y, X, device)["in"]

    def gather_windows(self, volume: Tensor, starts_dev: Tensor, n: int, roi) -> None:
        """Sliding-window gather of n windows of `volume` [C, VZ, VY, VX]: the blocked 16-bit copy feeds encoder1, the
        fp32 batch feeds the patch embedding (which reads the image itself)."""
        b = self._buffers(n, roi[0], roi[1], roi[2], volume.device)
        K.swi_gather(volume, starts_dev, n, roi, b["in"])
        K.swi_gather_ncdhw(volume, starts_dev, n, roi, b["x32"])

    # ---------------------------------------------------------------- pieces
    def _gemm(self, src: Blocked, segs, pw: PackedConv, dst, f32: bool, dst_cbt: int, dst_c0: int = 0) -> None:
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], self.nm)
        t = dst if isinstance(dst, torch.Tensor) else dst.t
        K.conv3d(src, pw, a_cb, t, _lib.OUT_BLOCKED_F32 if f32 else _lib.OUT_BLOCKED_BF16, dst_cbt=dst_cbt,
                 dst_cb_off=dst_c0 // 8)
        self.launches += 1

    def _conv_stats(self, src: Blocked, segs, pw: PackedConv, tag: str):
        """conv -> fp32 raw output + InstanceNorm (mean, rstd) table."""
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        cout = pw.n_out
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], self.nm)
        raw = self._ws.get("raw_" + tag, n * cout * Z * Y * X, torch.float32)
        tile = K.plan_conv_norm((X, Y, Z), n, pw, True, a_cb)
        stats = self._ws.get("stats", n * tile.tiles_per_img * cout * 2, torch.float32)
        K.conv3d(src, pw, a_cb, raw, _lib.OUT_BLOCKED_F32, stats=stats, dst_cbt=cout // 8, tile=tile)
        mr = self._ws.get("mr_" + tag, n * cout * 2, torch.float32)
        K.instnorm_finalize(stats, n, tile.tiles_per_img, cout, Z * Y * X, mr)
        self.launches += 2
        return raw, mr

    def _res_block(self, name: str, P, src: Blocked, segs, mid: Blocked, dst: Blocked, dst_c0: int) -> None:
        """MONAI UnetResBlock: conv1-IN-lrelu-conv2-IN, (+ conv3-IN | + x), lrelu."""
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        cout = P[name + ".c2"].n_out
        raw, mr = self._conv_stats(src, segs, P[name + ".c1"], "a")
        K.instnorm_act_apply(raw, True, mr, n, cout, Z, Y, X, mid, 0, LRELU_SLOPE)
        raw2, mr2 = self._conv_stats(mid, [(0, cout)], P[name + ".c2"], "b")
        if name + ".c3" in P:
            raw3, mr3 = self._conv_stats(src, segs, P[name + ".c3"], "a")
            K.instnorm_residual_act(raw2, True, mr2, raw3, True, mr3, cout // 8, 0, dst, dst_c0, n, cout, Z * Y * X, LRELU_SLOPE)
        else:
            assert len(segs) == 1 and segs[0][1] == cout
            K.instnorm_residual_act(raw2, True, mr2, src.t, False, None, src.cbt, segs[0][0], dst, dst_c0, n, cout, Z * Y * X,
                                    LRELU_SLOPE)
        self.launches += 2

    def _identity_table(self, n: int, channels: int, device) -> Tensor:
        key = (n, channels)
        t = self._ident.get(key)
        if t is None:
            t = torch.zeros((n, channels, 2), dtype=torch.float32, device=device)
            t[:, :, 1] = 1.0
            self._ident[key] = t
        return t

    # ---------------------------------------------------------------- forward
    @torch.no_grad()
    def forward_blocked(self, n: int, Z: int, Y: int, X: int, logits: Optional[Tensor], device=None, x: Optional[Tensor] = None):
        """x: the NCDHW fp32 input (the patch embedding reads it directly; the blocked copy in b['in'] feeds encoder1);
        None = the window batch gather_windows() left in the engine.  logits None: returns the last feature map (the
        caller fuses the 1x1x1 head into its consumer)."""
        _lib.require_device()
        net = self.net
        F = net.feature_size
        device = x.device if x is not None else (logits.device if logits is not None else device)
        P = self._pack()
        b = self._buffers(n, Z, Y, X, device)
        if x is None:
            x = b["x32"]
        if self._ws is None:
            self._ws = _Workspace(device)
        vit = net.swinViT
        win = tuple(net.window_size)
        dims = [(Z >> l, Y >> l, X >> l) for l in range(6)]
        vox = [d[0] * d[1] * d[2] for d in dims]

        # ---- SwinTransformer (MONAI swin_unetr.py SwinTransformer.forward)
        pe = vit.patch_embed.proj
        K.swin_patch_embed(x, pe.weight.detach(), pe.bias.detach() if pe.bias is not None else None, b["xs0"])
        hid_dst = {0: (b["hid0"], 0), 1: (b["hid1"], 0), 2: (b["hid2"], 0), 3: (b["cat5"], 8 * F), 4: (b["hid4"], 0)}
        if net.normalize:
            K.swin_layernorm(b["xs0"], n, F, vox[1], hid_dst[0][0], hid_dst[0][1])
        else:
            raise NotImplementedError("SwinUNETR(normalize=False) is not built")
        self.launches += 2
        for s in range(4):
            C_, d, nv = F << s, dims[s + 1], vox[s + 1]
            layer = getattr(vit, f"layers{s + 1}")[0]
            xs, ln, qkv, att, y, h32, h = (b[f"xs{s}"], b[f"ln{s}"], b[f"qkv{s}"], b[f"att{s}"], b[f"y{s}"], b[f"h32_{s}"],
                                           b[f"h{s}"])
            heads = layer.blocks[0].attn.num_heads
            depth = len(layer.blocks)
            blk0 = layer.blocks[0]
            K.swin_layernorm(xs, n, C_, nv, ln, 0, gamma=blk0.norm1.weight.detach(), beta=blk0.norm1.bias.detach(),
                             eps=blk0.norm1.eps)
            self.launches += 1
            for j, blk in enumerate(layer.blocks):
                p = f"l{s}.b{j}."
                shift = (0, 0, 0) if j % 2 == 0 else tuple(w // 2 for w in win)
                self._gemm(ln, [(0, C_)], P[p + "qkv"], qkv, False, qkv.cbt)
                K.swin_window_attention(qkv, att, blk.attn.relative_position_bias_table.detach(),
                                        blk.attn.qkv.bias.detach() if blk.attn.qkv.bias is not None else None, heads, win, shift)
                self._gemm(att, [(0, C_)], P[p + "proj"], y, True, C_ // 8)
                K.swin_layernorm(xs, n, C_, nv, ln, 0, add=y, gamma=blk.norm2.weight.detach(), beta=blk.norm2.bias.detach(),
                                 eps=blk.norm2.eps)
                self._gemm(ln, [(0, C_)], P[p + "fc1"], h32, True, 4 * C_ // 8)
                K.instnorm_act_apply(h32, True, self._identity_table(n, 4 * C_, device), n, 4 * C_, d[0], d[1], d[2], h, 0,
                                     gelu=True)
                self._gemm(h, [(0, 4 * C_)], P[p + "fc2"], y, True, C_ // 8)
                if j + 1 < depth:
                    nxt = layer.blocks[j + 1]
                    K.swin_layernorm(xs, n, C_, nv, ln, 0, add=y, gamma=nxt.norm1.weight.detach(), beta=nxt.norm1.bias.detach(),
                                     eps=nxt.norm1.eps)
                else:
                    K.swin_layernorm(xs, n, C_, nv, None, 0, add=y)
                self.launches += 4
            ds = layer.downsample
            K.swin_merge_ln(xs, n, C_, d[0], d[1], d[2], ds.norm.weight.detach(), ds.norm.bias.detach(), b[f"mg{s}"], ds.norm.eps)
            nxs = b[f"xs{s + 1}"]
            self._gemm(b[f"mg{s}"], [(0, 8 * C_)], P[f"l{s}.red"], nxs, True, 2 * C_ // 8)
            K.swin_layernorm(nxs, n, 2 * C_, vox[s + 2], hid_dst[s + 1][0], hid_dst[s + 1][1])
            self.launches += 2

        # ---- UNETR encoder / decoder (MONAI SwinUNETR.forward)
        self._res_block("enc1", P, b["in"], [(0, net.in_channels)], b["mid_d1"], b["cat1"], F)
        self._res_block("enc2", P, b["hid0"], [(0, F)], b["mid_d2"], b["cat2"], F)
        self._res_block("enc3", P, b["hid1"], [(0, 2 * F)], b["mid_d3"], b["cat3"], 2 * F)
        self._res_block("enc4", P, b["hid2"], [(0, 4 * F)], b["mid_d4"], b["cat4"], 4 * F)
        self._res_block("enc10", P, b["hid4"], [(0, 16 * F)], b["mid_e10"], b["e10"], 0)
        cur, cc = b["e10"], 16 * F
        for name, cat, c, out in (("dec5", "cat5", 8 * F, "d5"), ("dec4", "cat4", 4 * F, "d4"), ("dec3", "cat3", 2 * F, "d3"),
                                  ("dec2", "cat2", F, "d2"), ("dec1", "cat1", F, "d1")):
            a_cb = K.a_chunk_table(cur, [0], [cc], self.nm)
            K.conv3d(cur, P[name + ".up"], a_cb, b[cat].t, _lib.OUT_CONVT_K2S2, dst_cbt=b[cat].cbt, dst_cb_off=0)
            self.launches += 1
            self._res_block(name, P, b[cat], [(0, c), (c, c)], b["mid_" + out], b[out], 0)
            cur, cc = b[out], c
        if logits is None:
            return cur
        oc = net.out.conv.conv
        K.conv1x1_logits(cur, 0, F, oc.weight, oc.bias, logits)
        self.launches += 1
        return logits

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        _lib.require_device()
        if not x.is_cuda:
            raise RuntimeError("mmseg_b200 engines run on CUDA tensors only (no CPU fallback)")
        x = x.contiguous().float()
        n, cin, Z, Y, X = x.shape
        if cin != self.net.in_channels:
            raise ValueError(f"expected {self.net.in_channels} input channels, got {cin}")
        K.pack_ncdhw(x, self.input_buffer(n, Z, Y, X, x.device))
        logits = torch.empty((n, self.net.out_channels, Z, Y, X), dtype=torch.float32, device=x.device)
        return self.forward_blocked(n, Z, Y, X, logits, x=x)

    def hidden_states(self, n: int, Z: int, Y: int, X: int, device) -> List[Tensor]:
        """The five swinViT outputs of the last forward as NCDHW fp32 (reference swin_unetr.py:129-130)."""
        b = self._buffers(n, Z, Y, X, device)
        F = self.net.feature_size
        return [b["hid0"].to_ncdhw(), b["hid1"].to_ncdhw(), b["hid2"].to_ncdhw(), b["cat5"].to_ncdhw(8 * F, 8 * F),
                b["hid4"].to_ncdhw()]
