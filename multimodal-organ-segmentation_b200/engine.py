"""Forward engines: the launch sequences that evaluate UNet3D / DualEncoder with the sm_100a kernels.

An engine owns (a) kernel-layout copies of the module's weights (derived caches, rebuilt when a parameter changes)
and (b) the activation workspaces for one (n_img, Z, Y, X) problem size.  It issues only C-ABI kernel launches on
the current stream, so a whole forward can be captured in a CUDA graph.

Numeric modes (numerics.py — the parity ladder bf16 / fp16 / fp16w2 / fp16a2 / parity = bf16x3 / fp16x3)
  * "bf16"   — bf16 operands, fp32 accumulate (north_star's throughput mode); raw conv outputs stored bf16.
  * "fp16"   — the same single pass with fp16 operands and storage (11-bit significand, same MMA rate).
  * "parity" — 3-pass split-bf16 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulate), activations stored as
               bf16 hi+lo pairs and raw conv outputs as fp32: ~2^-16 relative operand error, which is what the
               stated logit/label tolerances need.
"""
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import kernels as K
from .kernels import Blocked, PackedConv
from .numerics import mode as numeric_mode

Tensor = torch.Tensor


class _Workspace:
    """Named scratch tensors that keep their address across calls (graph-capture safe)."""

    def __init__(self, device):
        self.device = device
        self._t: Dict[str, Tensor] = {}

    def get(self, name: str, numel: int, dtype) -> Tensor:
        t = self._t.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, device=self.device)
            self._t[name] = t
        return t


def norm_kind(norm) -> str:
    """"instance" | "group" | "batch" | "none" for a ConvBlock3D norm module (reference unet.py:29-41)."""
    if norm is None or isinstance(norm, torch.nn.InstanceNorm3d):
        if norm is not None and (norm.affine or norm.track_running_stats):
            raise NotImplementedError("InstanceNorm3d with affine / running statistics is never built by the reference")
        return "instance"
    if isinstance(norm, torch.nn.GroupNorm):
        return "group"
    if isinstance(norm, torch.nn.BatchNorm3d):
        if norm.training or not norm.track_running_stats:
            raise NotImplementedError("BatchNorm3d with batch statistics (train mode) is not built in the sm_100a path: "
                                      "call model.eval() (inference uses the running statistics)")
        return "batch"
    if isinstance(norm, torch.nn.Identity):
        return "none"
    raise NotImplementedError(f"norm module {type(norm).__name__} has no sm_100a kernel")


def norm_kind_train(norm) -> str:
    """norm_kind for the TRAINING path: BatchNorm3d in train mode is allowed there (batch statistics)."""
    if isinstance(norm, torch.nn.BatchNorm3d):
        return "batch"
    return norm_kind(norm)


class ConvRunner:
    """conv (+ InstanceNorm statistics) -> finalize -> normalise/activate(/pool), on blocked buffers."""

    def __init__(self, split, device):
        """split: NumericMode / mode name, or the legacy bool (False = bf16, True = parity)."""
        self.nm = numeric_mode(split)
        self.split = self.nm      # what Blocked / pack_conv_weight / a_chunk_table take
        self.device = device
        self.ws = _Workspace(device)
        self.launches = 0
        self._tables: Dict[Tuple, Tuple] = {}

    # K segments: list of (first channel in src, real channels)
    def conv_norm_act(self, src: Blocked, segs: Sequence[Tuple[int, int]], pw: PackedConv, dst: Blocked, dst_c0: int = 0,
                      pooled: Optional[Blocked] = None, pooled_c0: int = 0, slope: float = 0.0, tag: str = "",
                      norm=None, gelu: bool = False) -> None:
        """norm: the block's norm module (reference unet.py:29-41) — None / nn.InstanceNorm3d(affine=False): statistics
        from the conv epilogue; nn.GroupNorm: group statistics from the same partials + affine; nn.BatchNorm3d in eval
        mode: running statistics + affine (no statistics pass at all); nn.Identity: conv + bias + activation.  For the
        last three `pw` must carry the conv bias (it is only cancelled by InstanceNorm)."""
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], pw.nm or self.split)
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        cout = pw.n_out
        raw_f32 = (pw.nm or self.nm).raw_f32    # per layer in the mixed modes
        raw = self.ws.get("raw32" if raw_f32 else "raw16", n * cout * Z * Y * X, torch.float32 if raw_f32 else self.nm.dtype)
        tile = K.plan_conv_norm((X, Y, Z), n, pw, raw_f32, a_cb)
        kind = norm_kind(norm)
        out_mode = _lib.OUT_BLOCKED_F32 if raw_f32 else _lib.OUT_BLOCKED_BF16
        if kind in ("batch", "none"):
            K.conv3d(src, pw, a_cb, raw, out_mode, dst_cbt=cout // 8, tile=tile)
            mr, shift = self._static_table(norm, kind, n, cout, raw.device)
            K.instnorm_act_apply(raw, raw_f32, mr, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, pooled_c0, shift=shift, gelu=gelu)
            self.launches += 2
            return
        stats = self.ws.get("stats", n * tile.tiles_per_img * cout * 2, torch.float32)
        K.conv3d(src, pw, a_cb, raw, out_mode, stats=stats, dst_cbt=cout // 8, tile=tile)
        if kind == "group":
            mr = self.ws.get("mean_rstd", n * cout * 2, torch.float32)
            shift = self.ws.get("shift", n * cout, torch.float32)
            ga = norm.weight.detach().float() if norm.weight is not None else None
            be = norm.bias.detach().float() if norm.bias is not None else None
            K.groupnorm_finalize(stats, n, tile.tiles_per_img, cout, norm.num_groups, Z * Y * X, ga, be, mr, shift, norm.eps)
            K.instnorm_act_apply(raw, raw_f32, mr, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, pooled_c0, shift=shift, gelu=gelu)
            self.launches += 3
        elif tile.tiles_per_img <= 64:
            # InstanceNorm statistics finalized inside the apply kernel's prologue (no separate ~9 us launch) for the deep,
            # latency-bound levels; with more partial rows the per-block prologue (rows x 64 B from L2) shows up in the
            # HBM-bound apply kernels of the 96^3 / 48^3 levels (measured: 5.2 -> 4.5 TB/s), so those keep the launch
            K.instnorm_act_apply(raw, raw_f32, None, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, pooled_c0,
                                 stats=stats, tiles_per_img=tile.tiles_per_img, gelu=gelu)
            self.launches += 2
        else:
            mr = self.ws.get("mean_rstd", n * cout * 2, torch.float32)
            K.instnorm_finalize(stats, n, tile.tiles_per_img, cout, Z * Y * X, mr)
            K.instnorm_act_apply(raw, raw_f32, mr, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, pooled_c0, gelu=gelu)
            self.launches += 3

    def _static_table(self, norm, kind: str, n: int, cout: int, device):
        """(mean, rstd) / shift tables of the norms that need no statistics of the current input: BatchNorm3d in eval mode
        (running statistics; y = (x - rm) * gamma / sqrt(rv + eps) + beta) and Identity.  A few hundred floats of
        parameter preparation, cached per parameter version like the packed weights."""
        if kind == "none":
            key, ver = ("none", n, cout), 0
        else:
            tens = [norm.running_mean, norm.running_var] + ([norm.weight, norm.bias] if norm.affine else [])
            key, ver = (id(norm), n, cout), _param_version(tens)
        hit = self._tables.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        mr = torch.zeros((n, cout, 2), dtype=torch.float32, device=device)
        shift = None
        if kind == "none":
            mr[:, :, 1] = 1.0
        else:
            c = norm.num_features
            g = norm.weight.detach().float() if norm.affine else torch.ones(c, device=device)
            mr[:, :c, 0] = norm.running_mean.detach().float()
            mr[:, :c, 1] = g / torch.sqrt(norm.running_var.detach().float() + norm.eps)
            shift = torch.zeros((n, cout), dtype=torch.float32, device=device)
            if norm.affine:
                shift[:, :c] = norm.bias.detach().float()
        self._tables[key] = (ver, mr, shift)
        return mr, shift

    def conv_transpose(self, src: Blocked, segs, pw: PackedConv, dst: Blocked, dst_c0: int = 0) -> None:
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], pw.nm or self.split)
        K.conv3d(src, pw, a_cb, dst.t, _lib.OUT_CONVT_K2S2, dst_cbt=dst.cbt, dst_cb_off=dst_c0 // 8,
                 dst_lo_off=dst.lo_off)
        self.launches += 1

    def conv_act(self, src: Blocked, segs, pw: PackedConv, dst: Blocked, dst_c0: int = 0) -> None:
        """conv + bias straight to an activation buffer (no norm): 1x1 fusion projections."""
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], pw.nm or self.split)
        K.conv3d(src, pw, a_cb, dst.t, _lib.OUT_BLOCKED_BF16_HILO if dst.split else _lib.OUT_BLOCKED_BF16,
                 dst_cbt=dst.cbt, dst_cb_off=dst_c0 // 8, dst_lo_off=dst.lo_off)
        self.launches += 1

    def conv_logits(self, src: Blocked, segs, pw: PackedConv, out: Tensor, conv=None) -> None:
        if conv is not None and len(segs) == 1 and segs[0][0] % 8 == 0 and segs[0][1] % 8 == 0 \
                and conv.out_channels <= 16 and segs[0][1] <= 256:
            # thin 1x1 (K = 32, N = 8): CUDA cores at HBM speed instead of a 16-column tensor-core tile
            K.conv1x1_logits(src, segs[0][0], segs[0][1], conv.weight, conv.bias, out)
            self.launches += 1
            return
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], pw.nm or self.split)
        K.conv3d(src, pw, a_cb, out, _lib.OUT_NCDHW_F32)
        self.launches += 1


def _param_version(params: Sequence[Tensor]) -> Tuple:
    return tuple((p.data_ptr(), p._version) for p in params)


def module_version(module) -> Tuple:
    """Version stamp of everything a captured forward bakes in: parameters and norm buffers (BatchNorm running
    statistics).  In-place updates (optimizer.step, load_state_dict) bump `_version`; re-assignment changes data_ptr."""
    return _param_version(list(module.parameters()) + list(module.buffers()))


class UNet3DEngine:
    """UNet3D.forward (reference src/models/backbones/unet.py:165-200) on blocked buffers.

    `module` is the drop-in UNet3D nn.Module (same attribute tree / state_dict as the reference); only its
    parameters are read.  Dropout is applied by the caller (it is the identity in eval / p=0).
    """

    def __init__(self, module, mode: str = "bf16", weights_from: Optional["UNet3DEngine"] = None):
        """weights_from: another engine of the same module and mode whose packed weights this one reads (the extra
        batch slots of the sliding-window inferer: one kernel-layout copy of the weights, not one per slot)."""
        self.nm = numeric_mode(mode)
        self.module = module
        self.mode = mode
        self.split = self.nm      # what Blocked / pack_conv_weight / a_chunk_table take
        self._weights_from = weights_from
        self._packed: Optional[Dict[str, PackedConv]] = None
        self._packed_version = None
        self._bufs: Dict[Tuple, Dict[str, object]] = {}
        self._runner: Optional[ConvRunner] = None

    # ---------------------------------------------------------------- weights
    def _pack(self) -> Dict[str, PackedConv]:
        if self._weights_from is not None:
            P = self._weights_from._pack()
            self._norms = self._weights_from._norms
            return P
        m = self.module
        params = list(m.parameters())
        ver = _param_version(params)
        if self._packed is not None and ver == self._packed_version:
            return self._packed
        f = m.features
        nm = self.nm
        P: Dict[str, PackedConv] = {}

        self._norms: Dict[str, object] = {}

        L = len(f)
        bsplit = lambda buf: nm.buffer(buf).a_split     # is that activation buffer stored hi + lo?

        def block(name: str, blk, segs1, tag: str, src1: str, src2: str):
            # bias of a conv that feeds InstanceNorm(affine=False) is cancelled exactly by the mean subtraction; every
            # other norm option (batch / group / none, model.backbone.norm) keeps it
            inst = isinstance(blk.norm1, torch.nn.InstanceNorm3d)
            # mixed modes: the layer's passes follow its input buffer (activation split) and its tag (weight split)
            P[name + ".conv1"] = K.pack_conv_weight(blk.conv1.weight, None if inst else blk.conv1.bias,
                                                    nm.layer(tag + ".1", bsplit(src1)), segs1, use_bias=not inst)
            P[name + ".conv2"] = K.pack_conv_weight(blk.conv2.weight, None if inst else blk.conv2.bias,
                                                    nm.layer(tag + ".2", bsplit(src2)), None, use_bias=not inst)
            self._norms[name + ".conv1"], self._norms[name + ".conv2"] = blk.norm1, blk.norm2

        block("init_conv", m.init_conv, [m.in_channels], "enc0", "in", "mid0")
        if self.in_packed:
            # hi + lo input AND split first-conv weights with a thin input: the three operand passes become ONE K chunk —
            # virtual input channels [hi | lo | hi] (written by pack_ncdhw / swi_gather) against [W_hi | W_hi | W_lo]
            blk = m.init_conv
            inst = isinstance(blk.norm1, torch.nn.InstanceNorm3d)
            w = blk.conv1.weight.detach().float()
            w_hi = w.to(nm.dtype).float()
            wv = torch.cat([w_hi, w_hi, w - w_hi], dim=1)
            P["init_conv.conv1"] = K.pack_conv_weight(wv, None if inst else blk.conv1.bias, self._in_packed_mode(),
                                                      [3 * m.in_channels], use_bias=not inst)
        for i, enc in enumerate(m.encoders):
            block(f"encoders.{i}", enc.conv, [f[i]], f"enc{i + 1}", f"pool{i + 1}", f"mid{i + 1}")
        for j, dec in enumerate(m.decoders):
            lvl = L - 2 - j
            up_src = "bott" if j == 0 else f"dec{lvl + 1}"
            P[f"decoders.{j}.up"] = K.pack_conv_weight(dec.up.weight, dec.up.bias, nm.layer("up", bsplit(up_src)), None,
                                                       transposed=True)
            block(f"decoders.{j}", dec.conv, [f[lvl], f[lvl]], f"dec{lvl}", f"cat{lvl}", f"mid{lvl}")
        P["out_conv"] = K.pack_conv_weight(m.out_conv.weight, m.out_conv.bias, nm.layer("up", bsplit("dec0")), None)
        self._packed, self._packed_version = P, ver
        return P

    @property
    def in_packed(self) -> bool:
        """First layer with the split passes packed into one K chunk (kernels.Blocked.packed_split)."""
        nm, m = self.nm, self.module
        return (nm.buffer("in").a_split and nm.layer("enc0.1", True).w_split and 3 * m.in_channels <= 16
                and os.environ.get("MMSEG_IN_PACKED", "1") == "1")

    def _in_packed_mode(self):
        from .numerics import _resolved
        return _resolved(self.nm, False, False, self.nm.layer("enc0.1", True).raw_f32)

    # ---------------------------------------------------------------- buffers
    def _buffers(self, n: int, Z: int, Y: int, X: int, device) -> Dict[str, object]:
        key = (n, Z, Y, X)
        b = self._bufs.get(key)
        if b is not None:
            return b
        f = self.module.features
        L = len(f)
        if min(Z, Y, X) >> (L - 1) < 1:
            raise ValueError(f"spatial size {(Z, Y, X)} is too small for {L} resolution levels")
        # sizes that are not divisible by 2^(L-1): MaxPool3d(2) floors, the ConvTranspose output (2 * floor) then differs from
        # the skip and UpBlock3D resizes it trilinearly (reference unet.py:108-109) — see forward_blocked
        sp = self.nm.buffer          # per-buffer storage mode (hi-only vs hi + lo; mixed modes differ per buffer)
        b = {"in": Blocked(n, (self.module.in_channels + 15) // 16 * 16, Z, Y, X, sp("in"), device)}
        if self.in_packed:
            b["in"] = Blocked(n, 16, Z, Y, X, self._in_packed_mode(), device)
            b["in"].packed_split = True
        b["in"].t.zero_()   # the sliding-window gather writes only the blocks with real channels
        for l in range(L):
            z, y, x = Z >> l, Y >> l, X >> l
            b[f"mid{l}"] = Blocked(n, f[l], z, y, x, sp(f"mid{l}"), device)            # ConvBlock3D conv1 output
            if l < L - 1:
                b[f"cat{l}"] = Blocked(n, 2 * f[l], z, y, x, sp(f"cat{l}"), device)     # [up | skip]
                b[f"dec{l}"] = Blocked(n, f[l], z, y, x, sp(f"dec{l}"), device)         # decoder block output
            else:
                b[f"bott"] = Blocked(n, f[l], z, y, x, sp("bott"), device)
            if l > 0:
                b[f"pool{l}"] = Blocked(n, f[l - 1], z, y, x, sp(f"pool{l}"), device)    # MaxPool3d(2) of level l-1
            if l < L - 1 and ((Z >> l) & 1 or (Y >> l) & 1 or (X >> l) & 1):
                # odd level: the up-sampled tensor (2 * floor) is produced here and resized into cat{l}
                b[f"up{l}"] = Blocked(n, f[l], 2 * (z >> 1), 2 * (y >> 1), 2 * (x >> 1), sp(f"cat{l}"), device)
        self._bufs[key] = b
        return b

    def input_buffer(self, n: int, Z: int, Y: int, X: int, device) -> Blocked:
        return self._buffers(n, Z, Y, X, device)["in"]

    def gather_windows(self, volume: Tensor, starts_dev: Tensor, n: int, roi) -> None:
        """Sliding-window gather of n windows of `volume` [C, VZ, VY, VX] straight into the blocked input buffer."""
        K.swi_gather(volume, starts_dev, n, roi, self.input_buffer(n, roi[0], roi[1], roi[2], volume.device))

    # ---------------------------------------------------------------- forward
    @torch.no_grad()
    def forward_blocked(self, n: int, Z: int, Y: int, X: int, logits: Optional[Tensor], device=None):
        """Runs the net on the engine's input buffer; writes NCDHW fp32 logits [n, out_channels, Z, Y, X]."""
        _lib.require_device()
        m = self.module
        f = m.features
        L = len(f)
        P = self._pack()
        N = self._norms
        device = logits.device if logits is not None else device
        b = self._buffers(n, Z, Y, X, device)
        if self._runner is None:
            self._runner = ConvRunner(self.split, device)
        r = self._runner
        cin_p = (m.in_channels + 15) // 16 * 16
        # encoder (unet.py:181-187); each block's output lands in the skip half of its level's concat buffer
        r.conv_norm_act(b["in"], [(0, (3 if self.in_packed else 1) * m.in_channels)], P["init_conv.conv1"], b["mid0"],
                        norm=N["init_conv.conv1"])
        for l in range(L):
            last = l == L - 1
            if l > 0:
                r.conv_norm_act(b[f"pool{l}"], [(0, f[l - 1])], P[f"encoders.{l - 1}.conv1"], b[f"mid{l}"],
                                norm=N[f"encoders.{l - 1}.conv1"])
                name2 = f"encoders.{l - 1}.conv2"
            else:
                name2 = "init_conv.conv2"
            if last:
                r.conv_norm_act(b[f"mid{l}"], [(0, f[l])], P[name2], b["bott"], norm=N[name2])
            elif f"up{l}" in b:     # odd extents: MaxPool3d(2) floors — its own kernel instead of the fused 2x2x2 cells
                r.conv_norm_act(b[f"mid{l}"], [(0, f[l])], P[name2], b[f"cat{l}"], dst_c0=f[l], norm=N[name2])
                K.maxpool3d_2(b[f"cat{l}"], b[f"pool{l + 1}"], f[l], src_c0=f[l])
                r.launches += 1
            else:
                r.conv_norm_act(b[f"mid{l}"], [(0, f[l])], P[name2], b[f"cat{l}"], dst_c0=f[l],
                                pooled=b[f"pool{l + 1}"], norm=N[name2])
        # decoder (unet.py:190-192): up -> cat([up, skip]) -> ConvBlock3D
        cur = b["bott"]
        for j in range(L - 1):
            l = L - 2 - j
            if f"up{l}" in b:
                # x.shape != skip.shape: F.interpolate(x, size=skip.shape[2:], mode="trilinear", align_corners=True)
                # (reference unet.py:108-109) between the ConvTranspose and the concat
                r.conv_transpose(cur, [(0, f[l + 1])], P[f"decoders.{j}.up"], b[f"up{l}"], dst_c0=0)
                c = b[f"cat{l}"]
                K.pack_ncdhw(K.trilinear_resize(b[f"up{l}"].to_ncdhw(), (c.Z, c.Y, c.X)), c, 0)
                r.launches += 3
            else:
                r.conv_transpose(cur, [(0, f[l + 1])], P[f"decoders.{j}.up"], b[f"cat{l}"], dst_c0=0)
            r.conv_norm_act(b[f"cat{l}"], [(0, f[l]), (f[l], f[l])], P[f"decoders.{j}.conv1"], b[f"mid{l}"],
                            norm=N[f"decoders.{j}.conv1"])
            r.conv_norm_act(b[f"mid{l}"], [(0, f[l])], P[f"decoders.{j}.conv2"], b[f"dec{l}"],
                            norm=N[f"decoders.{j}.conv2"])
            cur = b[f"dec{l}"]
        if logits is None:   # features only: the caller fuses out_conv into its consumer (sliding-window blend)
            return cur
        r.conv_logits(cur, [(0, f[0])], P["out_conv"], logits, m.out_conv)
        return logits

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        _lib.require_device()
        if not x.is_cuda:
            raise RuntimeError("mmseg_b200 engines run on CUDA tensors only (no CPU fallback)")
        x = x.contiguous().float()
        n, _, Z, Y, X = x.shape
        K.pack_ncdhw(x, self.input_buffer(n, Z, Y, X, x.device))
        logits = torch.empty((n, self.module.out_channels, Z, Y, X), dtype=torch.float32, device=x.device)
        return self.forward_blocked(n, Z, Y, X, logits)

    def feature_ncdhw(self, n, Z, Y, X, level: int) -> Tensor:
        """Encoder feature of `level` (the skip half of the concat buffer) as NCDHW fp32 — return_features."""
        f = self.module.features
        b = self._bufs[(n, Z, Y, X)]
        if level == len(f) - 1:
            return b["bott"].to_ncdhw()
        return b[f"cat{level}"].to_ncdhw(f[level], f[level])


class DualEncoderEngine:
    """DualEncoder.forward (reference src/models/backbones/dual_encoder.py:112-199) on blocked buffers.

    The M per-modality encoders write every level's output modality-major into ONE blocked "stack" buffer
    (channel m*C_l + c), so torch.stack / torch.cat of the reference cost nothing; the level fusion then reads it once:
      * mean (anything the reference does not special-case: 'early', 'late', 'cross_attention'), 'add':
        weighted sum kernel with a uniform weight;
      * 'attention' (CrossModalAttention, :207-254): channel means -> gate MLP + softmax -> weighted sum;
      * 'concat' (:179-182): a 1x1 conv (tcgen05 GEMM) whose K runs over all M*C_l stack channels.
    The fused feature lands directly in the skip half of that level's decoder concat buffer.
    """

    def __init__(self, module, mode: str = "bf16", weights_from: Optional["DualEncoderEngine"] = None):
        self.nm = numeric_mode(mode)
        self.module = module
        self.mode = mode
        self.split = self.nm      # what Blocked / pack_conv_weight / a_chunk_table take
        self._weights_from = weights_from
        self._packed = None
        self._packed_version = None
        self._bufs: Dict[Tuple, Dict[str, object]] = {}
        self._runner: Optional[ConvRunner] = None
        self.last_gate_weights: List[Tensor] = []

    def _pack(self) -> Dict[str, PackedConv]:
        if self._weights_from is not None:
            return self._weights_from._pack()
        m = self.module
        ver = _param_version(list(m.parameters()))
        if self._packed is not None and ver == self._packed_version:
            return self._packed
        f, nm, M = m.features, self.nm, m.num_modalities
        P: Dict[str, PackedConv] = {}

        L = len(f)
        bsplit = lambda buf: nm.buffer(buf).a_split

        def block(name, blk, segs1, tag, src1, src2):
            P[name + ".conv1"] = K.pack_conv_weight(blk.conv1.weight, None, nm.layer(tag + ".1", bsplit(src1)), segs1, use_bias=False)
            P[name + ".conv2"] = K.pack_conv_weight(blk.conv2.weight, None, nm.layer(tag + ".2", bsplit(src2)), None, use_bias=False)

        for i, enc in enumerate(m.encoders):
            block(f"enc{i}.init", enc["init_conv"], [m.in_channels_per_modality], "enc0", "in", "mid0")
            for l, blk in enumerate(enc["blocks"]):
                block(f"enc{i}.blocks.{l}", blk.conv, [f[l]], f"enc{l + 1}", f"pool{l + 1}", f"mid{l + 1}")
        if m.fusion_type == "concat":
            for l, proj in enumerate(m.fusion_proj):
                P[f"fusion_proj.{l}"] = K.pack_conv_weight(proj.weight, proj.bias, nm.layer("up", bsplit(f"stack{l}")), [f[l]] * M)
        for j, dec in enumerate(m.decoder):
            lvl = L - 2 - j
            up_src = "bott" if j == 0 else f"dec{lvl + 1}"
            P[f"decoder.{j}.up"] = K.pack_conv_weight(dec.up.weight, dec.up.bias, nm.layer("up", bsplit(up_src)), None, transposed=True)
            block(f"decoder.{j}", dec.conv, [f[lvl], f[lvl]], f"dec{lvl}", f"cat{lvl}", f"mid{lvl}")
        P["out_conv"] = K.pack_conv_weight(m.out_conv.weight, m.out_conv.bias, nm.layer("up", bsplit("dec0")), None)
        self._packed, self._packed_version = P, ver
        return P

    def _buffers(self, n, Z, Y, X, device):
        key = (n, Z, Y, X)
        b = self._bufs.get(key)
        if b is not None:
            return b
        m = self.module
        f, L, M, sp = m.features, len(m.features), m.num_modalities, self.nm.buffer
        if any(d % (1 << (L - 1)) for d in (Z, Y, X)):
            raise NotImplementedError(f"spatial size {(Z, Y, X)} is not divisible by {1 << (L - 1)} (trilinear resize "
                                      "branch of UpBlock3D, reference unet.py:108-109, is not implemented)")
        b = {}
        for i in range(M):   # one input buffer per modality (zeroed once: gathers / packs only write the real channels' blocks)
            b[f"in{i}"] = Blocked(n, (m.in_channels_per_modality + 15) // 16 * 16, Z, Y, X, sp("in"), device)
            b[f"in{i}"].t.zero_()
        for l in range(L):
            z, y, x = Z >> l, Y >> l, X >> l
            b[f"mid{l}"] = Blocked(n, f[l], z, y, x, sp(f"mid{l}"), device)
            b[f"stack{l}"] = Blocked(n, M * f[l], z, y, x, sp(f"stack{l}"), device)       # every modality's level-l output
            if l < L - 1:
                b[f"cat{l}"] = Blocked(n, 2 * f[l], z, y, x, sp(f"cat{l}"), device)     # [up | fused skip]
                b[f"dec{l}"] = Blocked(n, f[l], z, y, x, sp(f"dec{l}"), device)
            else:
                b["bott"] = Blocked(n, f[l], z, y, x, sp("bott"), device)
            if l > 0:
                b[f"pool{l}"] = Blocked(n, f[l - 1], z, y, x, sp(f"pool{l}"), device)
        self._bufs[key] = b
        return b

    def gather_windows(self, volume: Tensor, starts_dev: Tensor, n: int, roi) -> None:
        """Sliding-window gather: modality i's channels of every window go to that modality's input buffer."""
        m = self.module
        cpm = m.in_channels_per_modality
        b = self._buffers(n, roi[0], roi[1], roi[2], volume.device)
        for i in range(m.num_modalities):
            K.swi_gather(volume[i * cpm:(i + 1) * cpm], starts_dev, n, roi, b[f"in{i}"])

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        _lib.require_device()
        if not x.is_cuda:
            raise RuntimeError("mmseg_b200 engines run on CUDA tensors only (no CPU fallback)")
        m = self.module
        x = x.contiguous().float()
        n, cin, Z, Y, X = x.shape
        cpm = m.in_channels_per_modality
        assert cin == m.num_modalities * cpm, f"expected {m.num_modalities * cpm} input channels, got {cin}"
        b = self._buffers(n, Z, Y, X, x.device)
        for i in range(m.num_modalities):   # modality i reads x[:, i*cpm:(i+1)*cpm] (dual_encoder.py:131-133)
            K.pack_ncdhw(x[:, i * cpm:(i + 1) * cpm].contiguous(), b[f"in{i}"])
        logits = torch.empty((n, m.out_channels, Z, Y, X), dtype=torch.float32, device=x.device)
        return self.forward_blocked(n, Z, Y, X, logits)

    @torch.no_grad()
    def forward_blocked(self, n: int, Z: int, Y: int, X: int, logits: Optional[Tensor], device=None):
        """Runs the net on the per-modality input buffers; writes NCDHW fp32 logits [n, out_channels, Z, Y, X]."""
        _lib.require_device()
        m = self.module
        f, L, M = m.features, len(m.features), m.num_modalities
        cpm = m.in_channels_per_modality
        P = self._pack()
        device = logits.device if logits is not None else device
        b = self._buffers(n, Z, Y, X, device)
        if self._runner is None:
            self._runner = ConvRunner(self.split, device)
        r = self._runner
        # encoders (dual_encoder.py:131-144)
        for i in range(M):
            r.conv_norm_act(b[f"in{i}"], [(0, cpm)], P[f"enc{i}.init.conv1"], b["mid0"])
            for l in range(L):
                if l > 0:
                    r.conv_norm_act(b[f"pool{l}"], [(0, f[l - 1])], P[f"enc{i}.blocks.{l - 1}.conv1"], b[f"mid{l}"])
                    name2 = f"enc{i}.blocks.{l - 1}.conv2"
                else:
                    name2 = f"enc{i}.init.conv2"
                r.conv_norm_act(b[f"mid{l}"], [(0, f[l])], P[name2], b[f"stack{l}"], dst_c0=i * f[l],
                                pooled=b[f"pool{l + 1}"] if l < L - 1 else None)
        # level fusion (_fuse_features, :167-199) into the skip half of the concat buffers / the bottleneck
        self.last_gate_weights = []
        for l in range(L):
            dst, c0 = (b["bott"], 0) if l == L - 1 else (b[f"cat{l}"], f[l])
            st = b[f"stack{l}"]
            if m.fusion_type == "concat":
                r.conv_act(st, [(i * f[l], f[l]) for i in range(M)], P[f"fusion_proj.{l}"], dst, dst_c0=c0)
            elif m.fusion_type == "add":
                K.modality_combine(st, M, f[l], dst, c0, None, 1.0)
            elif m.fusion_type == "attention":
                att = m.fusion_layers[l].attention
                pooled = K.channel_mean(st, 0, M * f[l])
                w = K.gate_mlp(pooled, att[2].weight, att[2].bias, att[4].weight, att[4].bias)
                self.last_gate_weights.append(w)
                K.modality_combine(st, M, f[l], dst, c0, w)
            else:
                K.modality_combine(st, M, f[l], dst, c0, None, 1.0 / M)
        # shared decoder (:147-152)
        cur = b["bott"]
        for j in range(L - 1):
            l = L - 2 - j
            r.conv_transpose(cur, [(0, f[l + 1])], P[f"decoder.{j}.up"], b[f"cat{l}"], dst_c0=0)
            r.conv_norm_act(b[f"cat{l}"], [(0, f[l]), (f[l], f[l])], P[f"decoder.{j}.conv1"], b[f"mid{l}"])
            r.conv_norm_act(b[f"mid{l}"], [(0, f[l])], P[f"decoder.{j}.conv2"], b[f"dec{l}"])
            cur = b[f"dec{l}"]
        if logits is None:   # features only: the caller fuses out_conv into its consumer (sliding-window blend)
            return cur
        r.conv_logits(cur, [(0, f[0])], P["out_conv"], logits, m.out_conv)
        return logits

    def features_ncdhw(self, n, Z, Y, X):
        """(encoder_features[m][l], fused_features[l]) as NCDHW fp32 copies — the return_features dict."""
        m = self.module
        f, L, M = m.features, len(m.features), m.num_modalities
        b = self._bufs[(n, Z, Y, X)]
        enc = [[b[f"stack{l}"].to_ncdhw(i * f[l], f[l]) for l in range(L)] for i in range(M)]
        fused = [b["bott"].to_ncdhw() if l == L - 1 else b[f"cat{l}"].to_ncdhw(f[l], f[l]) for l in range(L)]
        return enc, fused
