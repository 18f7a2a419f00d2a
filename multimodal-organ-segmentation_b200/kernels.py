"""Torch-facing wrappers over the C ABI: blocked activation buffers, weight packing, and one function per kernel.

PyTorch only provides device memory and the current stream here; all arithmetic happens in libmmseg_b200.so.
"""
import ctypes as C
from dataclasses import dataclass
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import lib, check
import os

from .tiling import plan_conv, plan_roll, ConvTile, ROLL_FLAG, ROLL_KPAIR_FLAG
from .numerics import NumericMode, mode as numeric_mode

Tensor = torch.Tensor


# torch.cuda.current_stream() walks through device-index resolution, is_available() and os.getenv on every call —
# a quarter of the host time of a launch-bound step (tools/swin_host_prof.py); the raw-handle query is one C call
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> C.c_void_p:
    if _raw_stream is not None and _cur_device is not None:
        return C.c_void_p(_raw_stream(_cur_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


# kernels launched by each C-ABI entry point (bench.py reports the count as `gpu_launches`)
KERNELS_PER_CALL = {"mmseg_cross_attention_bwd": 3, "mmseg_adamw_multi": 2, "mmseg_gate_mlp_bwd": 2, "mmseg_channel_stats": 2, "mmseg_dicece_fwd": 2, "mmseg_channel_mean": 2, "mmseg_modality_dot": 2, "mmseg_tversky_fwd": 2}
LAUNCHES = [0]
# when a list, every C-ABI call is bracketed by CUDA events on the current stream: (name, info, ev0, ev1)
PROFILE: Optional[list] = None
_INFO = [None]


# MMSEG_NVTX=1: every C-ABI call is an NVTX range named after its entry point (+ the layer description when a per-kernel
# profile is running), so that nsys / ncu timelines read in the path's own vocabulary
NVTX = os.environ.get("MMSEG_NVTX", "0") == "1"


def _call(name: str, *args) -> None:
    if NVTX:
        info = _INFO[0]
        torch.cuda.nvtx.range_push(name if not info else f"{name} {info.get('layer', '')}")
        try:
            return _call_inner(name, *args)
        finally:
            torch.cuda.nvtx.range_pop()
    return _call_inner(name, *args)


def _call_inner(name: str, *args) -> None:
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if prof is not None:
        e1.record()
        prof.append((name, _INFO[0], e0, e1))
        _INFO[0] = None
    check(rc, name)
    LAUNCHES[0] += KERNELS_PER_CALL.get(name, 1)


# --------------------------------------------------------------------------------------------- blocked buffers
class Blocked:
    """A blocked activation buffer [n_img, cbt, Z, Y, X, 8] bf16.

    `cb` channel blocks of real data per numeric plane; in split ("parity") mode the buffer holds a hi plane
    (blocks [0, cb)) followed by a lo plane (blocks [cb, 2cb)), i.e. lo_off == cb.
    """

    def __init__(self, n_img: int, channels: int, Z: int, Y: int, X: int, split, device):
        """split: a NumericMode / mode name (numerics.py), or the legacy bool (False = bf16, True = parity)."""
        assert channels % 8 == 0
        nm = numeric_mode(split)
        self.n_img, self.channels, self.Z, self.Y, self.X = n_img, channels, Z, Y, X
        self.cb = channels // 8
        self.nm = nm
        self.fmt = nm.fmt                      # element format of the 16-bit storage (MMSEG_FMT_*)
        self.split = nm.a_split                # hi + lo planes
        self.cbt = self.cb * (2 if self.split else 1)
        self.lo_off = self.cb if self.split else 0
        self.t = torch.empty((n_img, self.cbt, Z, Y, X, 8), dtype=nm.dtype, device=device)
        # "packed split" input buffer: the producers (pack_ncdhw / swi_gather) write the virtual channels
        # [hi(C) | lo(C) | hi(C)] into this (non-split) buffer instead of hi / lo planes — see engine.py
        self.packed_split = False

    @property
    def nvox(self) -> int:
        return self.Z * self.Y * self.X

    def to_ncdhw(self, c0: int = 0, channels: Optional[int] = None) -> Tensor:
        """fp32 NCDHW copy of channels [c0, c0+channels) (hi+lo summed) via the unpack kernel."""
        channels = self.channels - c0 if channels is None else channels
        assert c0 % 8 == 0
        out = torch.empty((self.n_img, channels, self.Z, self.Y, self.X), dtype=torch.float32, device=self.t.device)
        _call("mmseg_unpack_ncdhw", _ptr(self.t), _ptr(out), self.n_img, channels, self.Z, self.Y, self.X,
                                     self.cbt, c0 // 8, self.lo_off, self.fmt, _stream())
        return out


def pack_ncdhw(x: Tensor, dst: Blocked, c0: int = 0) -> None:
    """NCDHW fp32 -> blocked (channels zero-padded up to a multiple of 16 so a K chunk is always whole)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    n, Cc, Z, Y, X = x.shape
    cb = ((Cc + 15) // 16) * 2
    assert (n, Z, Y, X) == (dst.n_img, dst.Z, dst.Y, dst.X) and c0 % 8 == 0 and c0 // 8 + cb <= dst.cb
    lo_off = dst.lo_off
    if dst.packed_split:
        assert 3 * Cc <= 8 * cb and not dst.split and c0 == 0
        lo_off = -1
    _call("mmseg_pack_ncdhw", _ptr(x), _ptr(dst.t), n, Cc, Z, Y, X, dst.cbt, c0 // 8, lo_off, cb, dst.fmt, _stream())


# --------------------------------------------------------------------------------------------- weight packing
@dataclass
class PackedConv:
    """Kernel-layout weights of one conv (a derived cache; never saved in a state_dict)."""
    w: Tensor                 # bf16 [n_ntiles, n_kchunks, taps, 2, NT, 8]
    bias: Optional[Tensor]    # fp32 [n_ntiles*NT] or None
    ksize: int
    cin: int                  # real input channels (sum of segments)
    n_out: int                # GEMM columns, padded to a multiple of 16
    out_channels: int         # real output channels (per tap for convT)
    NT: int
    n_ntiles: int
    n_kchunks: int            # including the extra passes of the split modes
    split: bool               # any operand split (more than one pass)
    is_convt: bool = False
    nm: Optional[NumericMode] = None


def _pack_gemm_weight(wmat: Tensor, NT: int) -> Tensor:
    """wmat [n_out, cin_p, taps] fp32 (cin_p multiple of 16) -> kernel layout (K-major SWIZZLE_NONE core matrices).

    taps == 1:  [n_ntiles, n_kc, 1, 2, NT, 8]
    taps == 27: [n_ntiles, n_kc, 9 (dy, dx), 2 (k half), 3*NT rows, 8] with row = (2 - dz)*NT + n, i.e. the z taps are
                stored dz-DESCENDING so that a contiguous row range is a range of consecutive output planes
                (conv_tc.cu folds dz into the MMA N dimension).
    """
    n_out, cin_p, taps = wmat.shape
    if taps == 1:
        v = wmat.reshape(n_out // NT, NT, cin_p // 16, 2, 8, 1)
        return v.permute(0, 2, 5, 3, 1, 4).contiguous()
    assert taps == 27
    v = wmat.reshape(n_out // NT, NT, cin_p // 16, 2, 8, 3, 3, 3)      # [nt, n, kc, half, k8, dz, dy, dx]
    v = v.flip(5)                                                       # dz descending
    v = v.permute(0, 2, 6, 7, 3, 5, 1, 4)                               # [nt, kc, dy, dx, half, dzr, n, k8]
    return v.reshape(n_out // NT, cin_p // 16, 9, 2, 3 * NT, 8).contiguous()


def pack_conv_weight(weight: Tensor, bias: Optional[Tensor], split: bool, seg_channels: Optional[Sequence[int]] = None,
                     use_bias: bool = True, nt_cap: int = 64, transposed: bool = False) -> PackedConv:
    """Repack an nn.Conv3d / nn.ConvTranspose3d(k2,s2) weight for conv_tc.cu.

    seg_channels: real channels of each input segment (concat order); each segment is zero-padded to a multiple
    of 16 channels so that a 16-channel K chunk never straddles two producers.
    Split modes append K chunks (numerics.py): a+w split [W_hi | W_hi | W_lo] pairs with [A_hi | A_lo | A_hi];
    w split only [W_hi | W_lo] with [A | A]; a split only [W | W] with [A_hi | A_lo].
    """
    nm = numeric_mode(split)
    dt = nm.dtype
    w = weight.detach().float()
    if transposed:  # [Cin, Cout, 2,2,2] -> GEMM columns n = (((dz*2+dy)*CB + cb)*2 + dx)*8 + j  (co = cb*8 + j)
        cin, cout = w.shape[0], w.shape[1]
        assert tuple(w.shape[2:]) == (2, 2, 2) and cout % 16 == 0
        cbn = cout // 8
        wm = w.reshape(cin, cbn, 8, 4, 2).permute(3, 1, 4, 2, 0).reshape(8 * cout, cin, 1)
        b = None if bias is None else bias.detach().float().view(1, cbn, 1, 8).expand(4, cbn, 2, 8).reshape(-1)
        ksize, out_channels = 1, cout
    else:
        cout, cin = w.shape[0], w.shape[1]
        ksize = w.shape[2]
        assert ksize in (1, 3) and tuple(w.shape[2:]) == (ksize,) * 3
        wm = w.reshape(cout, cin, ksize ** 3)
        b = None if bias is None else bias.detach().float()
        out_channels = cout
    n_real = wm.shape[0]
    n_out = (n_real + 15) // 16 * 16
    seg = list(seg_channels) if seg_channels is not None else [cin]
    assert sum(seg) == cin
    parts, c = [], 0
    for s in seg:
        sp = (s + 15) // 16 * 16
        blk = torch.zeros((n_out, sp, wm.shape[2]), dtype=torch.float32, device=w.device)
        blk[:n_real, :s] = wm[:, c:c + s]
        parts.append(blk)
        c += s
    wp = torch.cat(parts, dim=1)
    NT = min(n_out, nt_cap)
    while n_out % NT:
        NT -= 16
    hi = wp.to(dt)
    p_hi = _pack_gemm_weight(hi.float(), NT).to(dt)
    parts_w = [p_hi] * (2 if nm.a_split else 1)
    if nm.w_split:
        lo = (wp - hi.float()).to(dt)
        parts_w.append(_pack_gemm_weight(lo.float(), NT).to(dt))
    packed = (torch.cat(parts_w, dim=1) if len(parts_w) > 1 else p_hi).contiguous()
    bias_p = None
    if use_bias and b is not None:
        bias_p = torch.zeros(n_out, dtype=torch.float32, device=w.device)
        bias_p[:n_real] = b
    return PackedConv(packed, bias_p, ksize, cin, n_out, out_channels, NT, n_out // NT, packed.shape[1], nm.passes > 1,
                      transposed, nm)


class PackPlan:
    """Kernel-side weight packing (mmseg_weights_repack): a persistent packed buffer plus the two int32 gather tables that
    describe where every GEMM element lives in the fp32 PyTorch-layout parameter.  `run()` re-derives the packed operand
    from the LIVE parameter with one launch (plus one tiny launch for a bias) — no ATen flip / permute / cat / copy, no
    allocation, graph-capturable, and the packed buffer keeps its address (captured graphs stay valid across updates).

    Forms (constructors below): `forward` (what pack_conv_weight builds, incl. concat segments, ConvTranspose and the hi / lo
    splits), `dgrad` (flipped taps, channels transposed), `convt_dgrad`, `k1_dgrad`.  pack_conv_weight remains the ATen
    restatement the tests compare this kernel against bit for bit."""

    def __init__(self, weight: Tensor, n_off: Sequence[int], k_off: Sequence[int], ksize: int, flip: bool, nm,
                 nt_cap: int, out_channels: int, cin: int, is_convt: bool = False, bias: Optional[Tensor] = None,
                 bias_idx: Optional[Sequence[int]] = None):
        nm = numeric_mode(nm)
        dev = weight.device
        assert weight.dtype == torch.float32 and weight.is_contiguous(), "parameters are fp32 contiguous"
        self.weight, self.bias_src, self.nm = weight, bias, nm
        n_out = len(n_off)
        assert n_out % 16 == 0 and len(k_off) % 16 == 0
        NT = min(n_out, nt_cap)
        while n_out % NT:
            NT -= 16
        self.n_kc = len(k_off) // 16
        self.hi_copies = 2 if nm.a_split else 1
        self.has_lo = bool(nm.w_split)
        n_kc_total = self.n_kc * (self.hi_copies + int(self.has_lo))
        taps2d, rows = (9, 3 * NT) if ksize == 3 else (1, NT)
        self.n_off = torch.tensor(list(n_off), dtype=torch.int32, device=dev)
        self.k_off = torch.tensor(list(k_off), dtype=torch.int32, device=dev)
        w = torch.empty((n_out // NT, n_kc_total, taps2d, 2, rows, 8), dtype=nm.dtype, device=dev)
        bias_p = None
        self.bias_idx = None
        if bias is not None:
            assert bias_idx is not None and len(bias_idx) == n_out and bias.dtype == torch.float32
            self.bias_idx = torch.tensor(list(bias_idx), dtype=torch.int32, device=dev)
            bias_p = torch.empty(n_out, dtype=torch.float32, device=dev)
        self.ksize, self.flip = ksize, bool(flip)
        self.pc = PackedConv(w, bias_p, ksize, cin, n_out, out_channels, NT, n_out // NT, n_kc_total, nm.passes > 1,
                             is_convt, nm)
        self.version = None

    def run(self) -> PackedConv:
        pc = self.pc
        _call("mmseg_weights_repack", _ptr(self.weight), _ptr(self.n_off), _ptr(self.k_off), _ptr(pc.w), pc.n_out, pc.NT,
              self.n_kc, pc.n_kchunks, self.ksize, 1 if self.flip else 0, self.hi_copies, 1 if self.has_lo else 0,
              self.nm.fmt, 1.0, _stream())
        if pc.bias is not None:
            _call("mmseg_gather_f32", _ptr(self.bias_src), _ptr(self.bias_idx), _ptr(pc.bias), pc.n_out, _stream())
        return pc

    def fresh(self) -> PackedConv:
        """run() only when the parameter changed since the last call (inference: weights are usually static)."""
        ver = (self.weight.data_ptr(), self.weight._version,
               None if self.bias_src is None else (self.bias_src.data_ptr(), self.bias_src._version))
        if ver != self.version:
            self.run()
            self.version = ver
        return self.pc

    # ---- constructors (same arguments as pack_conv_weight where they overlap)
    @staticmethod
    def _k_table(seg, elem_off):
        """K positions of concat segments (each padded to a whole 16-channel chunk) -> source offsets."""
        k_off, c = [], 0
        for s_ in seg:
            sp = (s_ + 15) // 16 * 16
            k_off += [elem_off(c + i) for i in range(s_)] + [-1] * (sp - s_)
            c += s_
        return k_off

    @classmethod
    def forward(cls, weight: Tensor, bias: Optional[Tensor], split, seg_channels: Optional[Sequence[int]] = None,
                use_bias: bool = True, nt_cap: int = 64, transposed: bool = False) -> "PackPlan":
        w = weight.detach()
        b = bias.detach() if (use_bias and bias is not None) else None
        if transposed:   # nn.ConvTranspose3d [Cin, Cout, 2,2,2]: column n = (((dz*2+dy)*CB + cb)*2 + dx)*8 + j
            cin, cout = w.shape[0], w.shape[1]
            assert tuple(w.shape[2:]) == (2, 2, 2) and cout % 16 == 0
            cbn = cout // 8
            n_off, bias_idx = [], []
            for zy in range(4):
                for cb in range(cbn):
                    for dx in range(2):
                        for j in range(8):
                            co = cb * 8 + j
                            n_off.append(co * 8 + zy * 2 + dx)
                            bias_idx.append(co)
            seg = list(seg_channels) if seg_channels is not None else [cin]
            k_off = cls._k_table(seg, lambda ci: ci * cout * 8)
            return cls(w, n_off, k_off, 1, False, split, nt_cap, cout, cin, True, b, bias_idx if b is not None else None)
        cout, cin, ksize = w.shape[0], w.shape[1], w.shape[2]
        assert ksize in (1, 3) and tuple(w.shape[2:]) == (ksize,) * 3
        taps = ksize ** 3
        n_pad = (cout + 15) // 16 * 16
        n_off = [n * cin * taps for n in range(cout)] + [-1] * (n_pad - cout)
        bias_idx = list(range(cout)) + [-1] * (n_pad - cout)
        seg = list(seg_channels) if seg_channels is not None else [cin]
        assert sum(seg) == cin
        k_off = cls._k_table(seg, lambda ci: ci * taps)
        return cls(w, n_off, k_off, ksize, False, split, nt_cap, cout, cin, False, b, bias_idx if b is not None else None)

    @classmethod
    def dgrad(cls, weight: Tensor, nt_cap: int = 64) -> "PackPlan":
        """The conv whose forward IS the input gradient of nn.Conv3d(k, padding k//2): weight.flip(2,3,4).transpose(0,1),
        i.e. GEMM column = input channel, K = output channel, taps mirrored."""
        w = weight.detach()
        cout, cin, ksize = w.shape[0], w.shape[1], w.shape[2]
        taps = ksize ** 3
        n_pad = (cin + 15) // 16 * 16
        n_off = [ci * taps for ci in range(cin)] + [-1] * (n_pad - cin)
        k_off = cls._k_table([cout], lambda co: co * cin * taps)
        return cls(w, n_off, k_off, ksize, True, False, nt_cap, cin, cout)

    @classmethod
    def convt_dgrad(cls, weight: Tensor, nt_cap: int = 64) -> "PackPlan":
        """Input gradient of ConvTranspose3d(k2,s2) through its k=1 GEMM view: column = input channel ci, K index =
        tap8*Cout + co (the channel order mmseg_unshuffle_k2s2 produces)."""
        w = weight.detach()
        cin, f = w.shape[0], w.shape[1]
        n_pad = (cin + 15) // 16 * 16
        n_off = [ci * f * 8 for ci in range(cin)] + [-1] * (n_pad - cin)
        k_off = cls._k_table([8 * f], lambda k: (k % f) * 8 + k // f)
        return cls(w, n_off, k_off, 1, False, False, nt_cap, cin, 8 * f)

    @classmethod
    def k1_dgrad(cls, weight: Tensor, nt_cap: int = 64) -> "PackPlan":
        """Input gradient of a 1x1x1 conv [Cout, Cin, 1,1,1]: column = ci, K = co."""
        w = weight.detach()
        cout, cin = w.shape[0], w.shape[1]
        n_pad = (cin + 15) // 16 * 16
        n_off = list(range(cin)) + [-1] * (n_pad - cin)
        k_off = cls._k_table([cout], lambda co: co * cin)
        return cls(w, n_off, k_off, 1, False, False, nt_cap, cin, cout)


class PackBatch:
    """mmseg_weights_repack_multi over a fixed list of PackPlans (same element format): one launch re-derives every packed
    operand of a model from its live fp32 parameters — what a training step does before its forward (the per-weight
    launches cost 0.6 ms of a 16.5 ms DualEncoder step and 1.5 ms of a SwinUNETR step)."""

    def __init__(self, plans: Sequence["PackPlan"]):
        self.plans = list(plans)
        assert self.plans and len({p.nm.fmt for p in self.plans}) == 1
        dev = self.plans[0].weight.device
        descs = (_lib.RepackDesc * len(self.plans))()
        block_desc: List[int] = []
        for i, p in enumerate(self.plans):
            pc, d = p.pc, descs[i]
            d.w, d.n_off, d.k_off, d.dst = p.weight.data_ptr(), p.n_off.data_ptr(), p.k_off.data_ptr(), pc.w.data_ptr()
            d.n_out, d.NT, d.n_kc, d.n_kc_total = pc.n_out, pc.NT, p.n_kc, pc.n_kchunks
            d.ksize, d.flip, d.hi_copies, d.has_lo = p.ksize, int(p.flip), p.hi_copies, int(p.has_lo)
            d.first_block = len(block_desc)
            total = pc.n_out * p.n_kc * 2 * p.ksize ** 3
            block_desc += [i] * ((total + 255) // 256)
        raw = bytes(descs)
        self.descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self.block_desc = torch.tensor(block_desc, dtype=torch.int32, device=dev)
        self.n_blocks = len(block_desc)
        self.ptrs = tuple(p.weight.data_ptr() for p in self.plans)

    def valid(self) -> bool:
        return all(p.weight.data_ptr() == q for p, q in zip(self.plans, self.ptrs))

    def run(self) -> None:
        _call("mmseg_weights_repack_multi", _ptr(self.descs), len(self.plans), _ptr(self.block_desc), self.n_blocks,
              self.plans[0].nm.fmt, 1.0, _stream())
        for p in self.plans:
            if p.pc.bias is not None:
                _call("mmseg_gather_f32", _ptr(p.bias_src), _ptr(p.bias_idx), _ptr(p.pc.bias), p.pc.n_out, _stream())


def a_chunk_table(src: Blocked, seg_c0: Sequence[int], seg_channels: Sequence[int], split) -> List[int]:
    """First channel block of every K chunk, in the order pack_conv_weight laid the chunks out."""
    nm = numeric_mode(split)
    hi = []
    for c0, s in zip(seg_c0, seg_channels):
        assert c0 % 16 == 0
        for j in range((s + 15) // 16):
            hi.append(c0 // 8 + 2 * j)
    assert getattr(src, "split", nm.a_split) == nm.a_split and getattr(src, "fmt", nm.fmt) == nm.fmt, \
        "activation buffer and numeric mode disagree"
    out = list(hi)
    if nm.a_split:
        out += [b + src.lo_off for b in hi]      # A_lo * W_hi
    if nm.w_split:
        out += hi                                 # A_hi * W_lo
    return out


def plan_conv_norm(src_dims, n_img: int, pw: "PackedConv", raw_f32: bool, a_cb: Optional[Sequence[int]] = None) -> ConvTile:
    """Tile plan of a conv whose output goes to InstanceNorm (raw blocked output + statistics): the rolling-z kernel
    for the k=3, C_out = 32 bf16 layers whose weights fit in shared memory, else the classic tile kernel."""
    X, Y, Z = src_dims
    # measured on B200 (96^3 x 8 windows): C_in 64: 0.639 vs 0.704 ms, C_in 32: 0.364 vs 0.371 ms, C_in <= 16 (one K chunk
    # per plane, the issue loop's per-plane work is not amortised): 0.227 vs 0.220 ms -> rolling-z from two K chunks up
    if (pw.ksize == 3 and pw.n_out == 32 and pw.NT == 32 and pw.bias is None and pw.n_kchunks >= 2
            and os.environ.get("MMSEG_NO_ROLL", "0") != "1"):
        # two adjacent K chunks per TMA stage when the channel blocks allow it (half the stage operations of the issue lane)
        kpb = 1
        if a_cb is not None and pw.n_kchunks % 2 == 0 and os.environ.get("MMSEG_ROLL_KPAIR", "1") == "1" \
                and all(a_cb[i + 1] == a_cb[i] + 2 for i in range(0, pw.n_kchunks, 2)):
            kpb = 2
        t = plan_roll(X, Y, Z, n_img, pw.n_kchunks, pw.n_out, kpb)
        if t is None and kpb == 2:
            t = plan_roll(X, Y, Z, n_img, pw.n_kchunks, pw.n_out, 1)
        if t is not None:
            return t
    return plan_conv(X, Y, Z, n_img, pw.n_kchunks, pw.n_out, pw.ksize, pw.NT)


# --------------------------------------------------------------------------------------------- kernels
def conv3d(src: Blocked, pw: PackedConv, a_cb: Sequence[int], dst: Tensor, out_mode: int, *,
           stats: Optional[Tensor] = None, dst_cbt: int = 0, dst_cb_off: int = 0, dst_lo_off: int = 0,
           tile: Optional[ConvTile] = None, flags: int = 0) -> ConvTile:
    """Launch mmseg_conv3d_fwd.  Returns the tile plan used (stats must hold tiles_per_img partial rows)."""
    _lib.require_device()
    assert len(a_cb) == pw.n_kchunks, (len(a_cb), pw.n_kchunks)
    if tile is None:
        tile = plan_conv(src.X, src.Y, src.Z, src.n_img, pw.n_kchunks, pw.n_out, pw.ksize, pw.NT)
    a = _lib.ConvArgs()
    a.src, a.weights, a.bias, a.dst = src.t.data_ptr(), pw.w.data_ptr(), (pw.bias.data_ptr() if pw.bias is not None else None), dst.data_ptr()
    a.stats_partial = stats.data_ptr() if stats is not None else None
    a.n_img, a.Z, a.Y, a.X = src.n_img, src.Z, src.Y, src.X
    a.src_cbt, a.ksize, a.n_kchunks = src.cbt, pw.ksize, pw.n_kchunks
    a.NT, a.n_ntiles = tile.NT, tile.n_ntiles
    a.TX, a.TY, a.TZ, a.stages = tile.TX, tile.TY, tile.TZ, tile.stages
    a.out_mode, a.out_channels = out_mode, pw.out_channels
    a.dst_cbt, a.dst_cb_off, a.dst_lo_off = dst_cbt, dst_cb_off, dst_lo_off
    assert pw.w.dtype == src.t.dtype, "weights and activations must share the 16-bit element format"
    a.flags = flags | (ROLL_FLAG if tile.roll else 0) | (ROLL_KPAIR_FLAG if (tile.roll and tile.kpb == 2) else 0) \
        | (_lib.CONV_FP16_FLAG if src.fmt == _lib.FMT_FP16 else 0)
    a.a_cb[:len(a_cb)] = list(a_cb)
    if stats is not None:
        need = src.n_img * tile.tiles_per_img * pw.n_out * 2
        assert stats.numel() >= need and stats.dtype == torch.float32
    if PROFILE is not None:
        n_real = pw.out_channels * (8 if pw.is_convt else 1)
        _INFO[0] = {"flops": 2.0 * src.n_img * src.nvox * pw.cin * n_real * pw.ksize ** 3,
                    "issued_flops": 2.0 * src.n_img * src.nvox * pw.n_kchunks * 16 * pw.n_out * pw.ksize ** 3,
                    "layer": f"k{pw.ksize} cin{pw.cin} n{n_real} {src.Z}x{src.Y}x{src.X} img{src.n_img}",
                    "tile": (tile.TX, tile.TY, tile.TZ, tile.NT, tile.stages),
                    "ctas": tile.tiles_per_img * src.n_img * tile.n_ntiles}
    _call("mmseg_conv3d_fwd", C.byref(a), _stream())
    return tile


def instnorm_finalize(stats: Tensor, n_img: int, tiles_per_img: int, channels: int, voxels: int, mean_rstd: Tensor,
                      eps: float = 1e-5) -> None:
    _call("mmseg_instnorm_finalize", _ptr(stats), n_img, tiles_per_img, channels, voxels, eps, _ptr(mean_rstd),
                                      _stream())


def groupnorm_finalize(stats: Tensor, n_img: int, tiles_per_img: int, channels: int, groups: int, voxels: int,
                       gamma: Optional[Tensor], beta: Optional[Tensor], mean_rstd: Tensor, shift: Tensor,
                       eps: float = 1e-5) -> None:
    _call("mmseg_groupnorm_finalize", _ptr(stats), n_img, tiles_per_img, channels, groups, voxels, eps,
          _ptr(gamma) if gamma is not None else None, _ptr(beta) if beta is not None else None, _ptr(mean_rstd),
          _ptr(shift), _stream())


def instnorm_act_apply(raw: Tensor, raw_is_f32: bool, mean_rstd: Optional[Tensor], n_img: int, channels: int,
                       Z: int, Y: int, X: int, dst: Blocked, dst_c0: int = 0, slope: float = 0.0,
                       pooled: Optional[Blocked] = None, pooled_c0: int = 0, *, stats: Optional[Tensor] = None,
                       tiles_per_img: int = 0, eps: float = 1e-5, mean_rstd_out: Optional[Tensor] = None,
                       shift: Optional[Tensor] = None, gelu: bool = False) -> None:
    """stats (the conv epilogue's partials) given: the statistics are finalized inside the apply kernel and
    mean_rstd is not read (no instnorm_finalize launch); mean_rstd_out optionally receives the table."""
    a = _lib.NormArgs()
    a.src, a.dst = raw.data_ptr(), dst.t.data_ptr()
    a.act = 1 if gelu else 0
    a.mean_rstd = mean_rstd.data_ptr() if mean_rstd is not None else None
    if shift is not None:
        assert stats is None and shift.dtype == torch.float32 and shift.numel() >= n_img * channels
        a.shift = shift.data_ptr()
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.numel() >= n_img * tiles_per_img * channels * 2 and tiles_per_img >= 1
        a.stats_partial, a.tiles_per_img, a.eps = stats.data_ptr(), tiles_per_img, eps
        a.mean_rstd_out = mean_rstd_out.data_ptr() if mean_rstd_out is not None else None
    a.pooled = pooled.t.data_ptr() if pooled is not None else None
    a.n_img, a.cb, a.Z, a.Y, a.X = n_img, channels // 8, Z, Y, X
    a.src_is_f32 = 1 if raw_is_f32 else 0
    a.dst_cbt, a.dst_cb_off, a.dst_lo_off = dst.cbt, dst_c0 // 8, dst.lo_off
    if pooled is not None:
        a.pool_cbt, a.pool_cb_off, a.pool_lo_off = pooled.cbt, pooled_c0 // 8, pooled.lo_off
    a.slope = slope
    a.elem_fmt = dst.fmt
    if PROFILE is not None:
        nel = n_img * channels * Z * Y * X
        out_b = 2 * (2 if dst.split else 1)
        _INFO[0] = {"bytes": nel * ((4 if raw_is_f32 else 2) + out_b) + (nel // 8 * out_b if pooled is not None else 0),
                    "layer": f"c{channels} {Z}x{Y}x{X} img{n_img} pool{int(pooled is not None)}"}
    _call("mmseg_instnorm_act_apply", C.byref(a), _stream())


def pack_ncdhw_ex(x: Tensor, dst: Blocked, c0: int = 0, pre_sigmoid: Optional[Tuple[float, float]] = None,
                  gate_logits: Optional[Tensor] = None) -> None:
    """pack_ncdhw with x <- sigmoid((x - a) * b) (pre_sigmoid = (a, b)) and / or x <- x * (1 + sigmoid(gate)) folded in
    (gate_logits [n_img, 1, Z, Y, X] fp32): the element-wise steps of SUVGuidedAttention."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    n, Cc, Z, Y, X = x.shape
    cb = (Cc + 15) // 16 * 2
    a, b = pre_sigmoid if pre_sigmoid is not None else (0.0, 1.0)
    if gate_logits is not None:
        assert gate_logits.dtype == torch.float32 and gate_logits.is_contiguous() and gate_logits.numel() == n * Z * Y * X
    _call("mmseg_pack_ncdhw_ex", _ptr(x), _ptr(dst.t), n, Cc, Z, Y, X, dst.cbt, c0 // 8, dst.lo_off, cb,
          1 if pre_sigmoid is not None else 0, float(a), float(b), _ptr(gate_logits) if gate_logits is not None else None,
          dst.fmt, _stream())


def swi_gather(volume: Tensor, starts_dev: Tensor, n_win: int, roi: Tuple[int, int, int], dst: Blocked) -> None:
    Cc, VZ, VY, VX = volume.shape
    # only the channel blocks that hold real channels are written: the zero padding up to a whole 16-channel K chunk
    # is written once when the engine allocates (and zeroes) its input buffer
    cb = (Cc + 7) // 8
    lo_off = dst.lo_off
    if dst.packed_split:
        cb, lo_off = (3 * Cc + 7) // 8, -1
    _call("mmseg_swi_gather", _ptr(volume), Cc, VZ, VY, VX, _ptr(starts_dev), n_win, roi[0], roi[1], roi[2],
                               _ptr(dst.t), dst.cbt, lo_off, cb, dst.fmt, _stream())


def swi_gather_ncdhw(volume: Tensor, starts_dev: Tensor, n_win: int, roi: Tuple[int, int, int], dst: Tensor) -> None:
    Cc, VZ, VY, VX = volume.shape
    assert dst.dtype == torch.float32 and dst.is_contiguous() and dst.numel() >= n_win * Cc * roi[0] * roi[1] * roi[2]
    _call("mmseg_swi_gather_ncdhw", _ptr(volume), Cc, VZ, VY, VX, _ptr(starts_dev), n_win, roi[0], roi[1], roi[2], _ptr(dst),
          _stream())


def swi_blend(win_logits: Tensor, starts_dev: Tensor, n_win: int, wz: Tensor, wy: Tensor, wx: Tensor, w_floor: float,
              out: Tensor, count: Tensor, box: Tuple[int, int, int, int, int, int]) -> None:
    K, VZ, VY, VX = out.shape
    _, _, RZ, RY, RX = win_logits.shape
    _call("mmseg_swi_blend", _ptr(win_logits), _ptr(starts_dev), n_win, K, RZ, RY, RX, _ptr(wz), _ptr(wy), _ptr(wx),
                              w_floor, _ptr(out), _ptr(count), VZ, VY, VX, *box, _stream())


def swi_logits_blend(feat: Blocked, c0: int, cin: int, window: int, weight: Tensor, bias: Optional[Tensor],
                     start_dev: Tensor, wz: Tensor, wy: Tensor, wx: Tensor, w_floor: float, out: Tensor, count: Tensor) -> None:
    """out_conv + blend of window `window` of the blocked feature batch `feat` in one kernel (no logits tensor)."""
    Kc, VZ, VY, VX = out.shape
    assert weight.shape[0] == Kc and weight.shape[1] == cin and c0 % 8 == 0
    w = weight.detach().reshape(Kc, cin)
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    b = bias.detach().float() if bias is not None else None
    _call("mmseg_swi_logits_blend", _ptr(feat.t), feat.cbt, c0 // 8, feat.lo_off if feat.split else 0, cin, window, _ptr(w),
          _ptr(b) if b is not None else None, Kc, _ptr(start_dev), feat.Z, feat.Y, feat.X, _ptr(wz), _ptr(wy), _ptr(wx),
          w_floor, _ptr(out), _ptr(count), VZ, VY, VX, feat.fmt, _stream())


def swi_finalize(out: Tensor, count: Tensor, normalize_in_place: bool, labels: Optional[Tensor], z0: int = 0,
                 z1: Optional[int] = None) -> None:
    """out [K, VZ, VY, VX] / count [VZ, VY, VX] (contiguous); finalizes the axis-0 slab [z0, z1) in place.
    labels (optional): uint8 [z1 - z0, VY, VX]."""
    K, VZ, VY, VX = out.shape
    z1 = VZ if z1 is None else z1
    assert out.is_contiguous() and count.is_contiguous() and 0 <= z0 < z1 <= VZ
    plane, row = VZ * VY * VX, VY * VX
    if labels is not None:
        assert labels.is_contiguous() and labels.numel() == (z1 - z0) * row and labels.dtype == torch.uint8
    _call("mmseg_swi_finalize", C.c_void_p(out.data_ptr() + 4 * z0 * row), C.c_void_p(count.data_ptr() + 4 * z0 * row), K,
          (z1 - z0) * row, plane, 1 if normalize_in_place else 0, _ptr(labels), _stream())


def swi_add_partial(acc: Tensor, z0: int, z1: int, part: Tensor) -> None:
    """acc[:, z0:z1] += part — acc [P, VZ, VY, VX] contiguous fp32, part [P, z1 - z0, VY, VX] contiguous fp32."""
    P, VZ, VY, VX = acc.shape
    row = VY * VX
    assert acc.is_contiguous() and part.is_contiguous() and tuple(part.shape) == (P, z1 - z0, VY, VX)
    assert acc.dtype == part.dtype == torch.float32
    _call("mmseg_swi_add_partial", C.c_void_p(acc.data_ptr() + 4 * z0 * row), VZ * row, _ptr(part), (z1 - z0) * row, P,
          (z1 - z0) * row, _stream())


def dicece_fwd(logits: Tensor, target: Tensor, dice_weight: float, ce_weight: float, smooth: float = 1.0,
               include_background: bool = True, class_weights: Optional[Tensor] = None):
    """Returns (result[3] = (total, dice, ce), sums[B, 3C+2]) on the device."""
    _lib.require_device()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    assert target.dtype == torch.int64 and target.is_contiguous()
    B, Cc = logits.shape[:2]
    N = logits[0, 0].numel()
    n_blocks = max(1, min(148 * 4, (N + 255) // 256))
    partial = torch.empty((B, n_blocks, 3 * Cc + 2), dtype=torch.float32, device=logits.device)
    result = torch.empty(3, dtype=torch.float32, device=logits.device)
    sums = torch.empty((B, 3 * Cc + 2), dtype=torch.float32, device=logits.device)
    _call("mmseg_dicece_fwd", _ptr(logits), _ptr(target), B, Cc, N, dice_weight, ce_weight, smooth,
                               1 if include_background else 0, _ptr(class_weights), _ptr(partial), n_blocks,
                               _ptr(result), _ptr(sums), _stream())
    return result, sums


def dicece_bwd(logits: Tensor, target: Tensor, sums: Tensor, grad_out: Optional[Tensor], dice_weight: float,
               ce_weight: float, smooth: float = 1.0, include_background: bool = True,
               class_weights: Optional[Tensor] = None) -> Tensor:
    B, Cc = logits.shape[:2]
    N = logits[0, 0].numel()
    dl = torch.empty_like(logits)
    go = None if grad_out is None else grad_out.reshape(1).float().contiguous()
    _call("mmseg_dicece_bwd", _ptr(logits), _ptr(target), B, Cc, N, dice_weight, ce_weight, smooth,
                               1 if include_background else 0, _ptr(class_weights), _ptr(sums), _ptr(go), _ptr(dl),
                               _stream())
    return dl


def channel_mean(src: Blocked, c0: int, channels: int) -> Tensor:
    """Global average pool per (image, channel) of channels [c0, c0+channels) -> fp32 [n_img, channels]."""
    cb = channels // 8
    n_chunks = max(1, min(64, (src.nvox + 4095) // 4096))
    partial = torch.empty((src.n_img * cb, n_chunks, 8), dtype=torch.float32, device=src.t.device)
    mean = torch.empty((src.n_img, channels), dtype=torch.float32, device=src.t.device)
    _call("mmseg_channel_mean", _ptr(src.t), src.n_img, src.cbt, c0 // 8, src.lo_off, cb, src.nvox, _ptr(partial),
                                 n_chunks, _ptr(mean), src.fmt, _stream())
    return mean


def gate_mlp(pooled: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    n, MC = pooled.shape
    H, M = w1.shape[0], w2.shape[0]
    out = torch.empty((n, M), dtype=torch.float32, device=pooled.device)
    f = lambda t: t.detach().float().contiguous()
    w1, b1, w2, b2 = f(w1), f(b1), f(w2), f(b2)
    _call("mmseg_gate_mlp", _ptr(pooled), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), n, MC, H, M, _ptr(out), _stream())
    return out


def gate_mlp_bwd(pooled: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, dweights: Tensor):
    """Backward of gate_mlp: returns (dpooled [n, MC], dW1, db1, dW2, db2) — fp32, parameter gradients summed over images."""
    n, MC = pooled.shape
    H, M = w1.shape[0], w2.shape[0]
    dev = pooled.device
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    ws, dpooled, dw1, db1, dw2, db2 = f((2 * H + M) * n), f(n, MC), f(H, MC), f(H), f(M, H), f(M)
    c32 = lambda t: t.detach().float().contiguous()
    _call("mmseg_gate_mlp_bwd", _ptr(c32(pooled)), _ptr(c32(w1)), _ptr(c32(b1)), _ptr(c32(w2)), _ptr(c32(b2)),
          _ptr(c32(dweights)), n, MC, H, M, _ptr(ws), _ptr(dpooled), _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _stream())
    return dpooled, dw1, db1, dw2, db2


def modality_combine(src: Blocked, M: int, channels: int, dst: Blocked, dst_c0: int, weights: Optional[Tensor],
                     uniform_weight: float = 1.0) -> None:
    _call("mmseg_modality_combine", _ptr(src.t), src.n_img, src.cbt, src.lo_off, M, channels // 8, src.nvox,
                                     _ptr(weights), uniform_weight, _ptr(dst.t), dst.cbt, dst_c0 // 8, dst.lo_off,
                                     dst.fmt, _stream())


def maxpool3d_2(src: Blocked, dst: Blocked, channels: Optional[int] = None, src_c0: int = 0, dst_c0: int = 0) -> None:
    channels = src.channels if channels is None else channels
    _call("mmseg_maxpool3d_2", _ptr(src.t), src.n_img, src.cbt, src_c0 // 8, src.lo_off, channels // 8, src.Z, src.Y,
                                src.X, _ptr(dst.t), dst.cbt, dst_c0 // 8, dst.lo_off, dst.fmt, _stream())


# --------------------------------------------------------------------------------------------- weight gradient
def _largest_divisor(cands, values):
    for c in cands:
        if all(v % c == 0 for v in values):
            return c
    raise ValueError(f"no channel group size in {cands} divides {values}")


def wgrad_groups(ksize: int, seg_pad: Sequence[int], cout_pad: int) -> Tuple[int, int]:
    """(input channels per accumulator-row group, output channels per MMA column group) of the wgrad kernel."""
    if os.environ.get("MMSEG_WGRAD_OLD_GROUPS", "0") == "1":
        return (_largest_divisor((32, 16) if ksize == 3 else (128, 64, 32, 16), seg_pad),
                _largest_divisor((32, 16) if ksize == 3 else (256, 128, 64, 32, 16), [cout_pad]))
    return (_largest_divisor((32, 24, 16) if ksize == 3 else (128, 96, 64, 48, 32, 16), seg_pad),
            _largest_divisor((48, 32, 16) if ksize == 3 else (256, 192, 128, 96, 64, 48, 32, 16), [cout_pad]))


@lru_cache(maxsize=None)
def _plan_wgrad_tile(X: int, Y: int, Z: int, ksize: int, cig_blocks: int, cot_blocks: int) -> Tuple[int, int, int]:
    """(TX, TY, TZ) accepted by the library: largest K rows per plane (TX*TY, multiple of 16) that fits shared memory."""
    best = None
    for nx in range(1, X + 1):
        TX = (X + nx - 1) // nx
        if TX > 128:
            continue
        for TY in range(min(Y, 32), 0, -1):
            if (TX * TY) % 16:
                continue
            a = _lib.WgradArgs()
            a.n_img, a.Z, a.Y, a.X, a.ksize = 1, Z, Y, X, ksize
            a.TX, a.TY, a.TZ = TX, TY, min(Z, 8)
            a.cig_blocks, a.cot_blocks, a.n_cig, a.n_cot = cig_blocks, cot_blocks, 1, 1
            a.x_cbt, a.y_cbt, a.y_cb0, a.n_part = cig_blocks, cot_blocks, 0, 1
            if lib.mmseg_conv3d_wgrad_smem_bytes(C.byref(a)) > 0:
                rows = TX * TY
                waste = ((X + TX - 1) // TX * TX) * ((Y + TY - 1) // TY * TY) / float(X * Y)
                score = (min(rows, 512) / waste, -nx)
                if best is None or score > best[0]:
                    best = (score, (TX, TY, min(Z, 8)))
                break
        if best is not None and nx >= 4:
            break
    if best is None:
        # rows must be a multiple of 16: pad the tile beyond the volume (TMA zero-fills, so the sum is unchanged)
        TX = min(X, 16) if X >= 16 else 16
        TY = max(1, 16 // TX) if TX * max(1, 16 // TX) % 16 == 0 else 16
        best = (None, (TX, TY, min(Z, 8)))
    return best[1]


@lru_cache(maxsize=None)
def _wgrad_plan(segs: Tuple[Tuple[int, int], ...], X: int, Y: int, Z: int, n_img: int, cout_gemm: int, ksize: int,
                tz_env: str, groups_env: str):
    """Host-side plan of one wgrad launch (cached per layer shape: the accumulator-row map alone is O(Cin) Python)."""
    seg_ch = [s[1] for s in segs]
    seg_pad = [(s + 15) // 16 * 16 for s in seg_ch]
    cin = sum(seg_ch)
    # channel groups: input channels per accumulator-row group (k3: 3 dx copies x cig <= 128 rows), output channels per
    # MMA column group (k3: 3 dz taps x ntc <= 256 columns).  The 24- / 48- / 96-channel sizes serve SwinUNETR's widths
    # (48 * 2^s): C = 48 gets (24, 48) = 72 of 128 rows and N = 144 instead of (16, 16) = 48 rows and N = 48.
    cig, ntc = wgrad_groups(ksize, seg_pad, (cout_gemm + 15) // 16 * 16)
    cout_pad = (cout_gemm + 15) // 16 * 16
    groups, ci_map = [], []
    for (c0, s), sp in zip(segs, seg_pad):
        assert c0 % 8 == 0
        for j in range(sp // cig):
            groups.append((c0 + j * cig) // 8)
        base = (len(groups) - sp // cig) * cig
        ci_map.extend(base + i for i in range(s))
    n_cig, n_cot = len(groups), cout_pad // ntc
    assert n_cig <= _lib.MAX_WGRAD_GROUPS
    TX, TY, TZ = _plan_wgrad_tile(X, Y, Z, ksize, cig // 8, ntc // 8)
    if ksize == 3 and tz_env != "8":
        # z extent of a tile: every tile issues the MMAs (and loads) of TZ + 2 input planes for TZ output planes, so a
        # longer z run cuts the halo overhead (TZ = 8: 25 %, 32: 6 %) as long as the persistent CTAs stay balanced
        ctas = max(1, 148 // (n_cig * n_cot))
        txy = -(-X // TX) * -(-Y // TY) * n_img
        best = None
        for tz in ([int(tz_env)] if tz_env != "auto" else (8, 12, 16, 24, 32, 48, 64, 96, 128)):
            if tz > max(Z, 8):
                continue
            tz = min(tz, Z)
            nt = txy * -(-Z // tz)
            planes = -(-nt // min(nt, ctas)) * (tz + 2)          # input planes swept by the busiest CTA
            if best is None or planes < best[0]:
                best = (planes, tz)
        TZ = best[1]
    n_tiles = -(-X // TX) * -(-Y // TY) * -(-Z // TZ) * n_img
    n_part = max(1, min(n_tiles, 148 // (n_cig * n_cot)))
    return cin, cig, ntc, tuple(groups), tuple(ci_map), n_cig, n_cot, TX, TY, TZ, n_part


_CI_MAPS: dict = {}
_IM2COL_WS: dict = {}


def _im2col_workspace(numel: int, like: Tensor) -> Tensor:
    """The 32-channel im2col tensor of a 1-channel first layer (268 MB at 2 x 128^3).  Eager steps reuse ONE buffer per
    (device, stream) — uses are ordered on that stream; a fresh transient block of this size per step fragments the
    side stream's allocator pool (measured: +4 ms per step from repeated cudaMalloc).  Inside a CUDA-graph capture the
    buffer comes from the graph's own pool."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(numel, dtype=like.dtype, device=like.device)
    key = (like.device, _stream().value, like.dtype)
    ws = _IM2COL_WS.get(key)
    if ws is None or ws.numel() < numel:
        if len(_IM2COL_WS) >= 4:      # streams come and go with their engines: do not pin a buffer per dead stream
            _IM2COL_WS.clear()        # (a dropped buffer returns to its stream's pool, where reuse is stream-ordered)
        ws = _IM2COL_WS[key] = torch.empty(numel, dtype=like.dtype, device=like.device)
    return ws[:numel]


def conv3d_wgrad(x: Blocked, segs: Sequence[Tuple[int, int]], dy: Tensor, dy_cbt: int, dy_cb0: int, cout_gemm: int,
                 ksize: int, weight_shape, transposed: bool = False) -> Tensor:
    """dW (fp32, PyTorch weight layout `weight_shape`) of a conv whose input is `x` (channel segments `segs` in concat
    order) and whose raw-output gradient is `dy` (blocked bf16, `dy_cbt` channel blocks per image, first block dy_cb0)."""
    _lib.require_device()
    assert not x.split and x.fmt == _lib.FMT_BF16, "the backward path runs in bf16 mode"
    segs_key = tuple((int(c0), int(s_)) for c0, s_ in segs)
    if ksize == 3 and not transposed and len(segs_key) == 1 and segs_key[0][1] == 1 \
            and os.environ.get("MMSEG_WGRAD_IM2COL", "1") == "1":
        # one input channel (DualEncoder's per-modality first layers): 3 live accumulator rows of 128 for the MMAs of a
        # 32-channel layer — the 27 shifted copies as a 32-channel tensor + the k = 1 wgrad GEMM is HBM-bound instead
        c0 = segs_key[0][0]
        col = Blocked.__new__(Blocked)
        col.n_img, col.channels, col.Z, col.Y, col.X, col.cb, col.cbt, col.lo_off = x.n_img, 32, x.Z, x.Y, x.X, 4, 4, 0
        col.nm, col.fmt, col.split, col.packed_split = getattr(x, "nm", None), x.fmt, False, False
        col.t = _im2col_workspace(x.n_img * 32 * x.nvox, x.t).view(x.n_img, 4, x.Z, x.Y, x.X, 8)
        _call("mmseg_im2col_k3_c1", _ptr(x.t), x.n_img, x.cbt, c0 // 8, c0 % 8, x.Z, x.Y, x.X, _ptr(col.t), _stream())
        dw = conv3d_wgrad(col, [(0, 27)], dy, dy_cbt, dy_cb0, cout_gemm, 1, (weight_shape[0], 27, 1, 1, 1))
        return dw.view(tuple(weight_shape))
    (cin, cig, ntc, groups, ci_map, n_cig, n_cot, TX, TY, TZ, n_part) = _wgrad_plan(
        segs_key, x.X, x.Y, x.Z, x.n_img, cout_gemm, ksize,
        os.environ.get("MMSEG_WGRAD_TZ", "auto"), os.environ.get("MMSEG_WGRAD_OLD_GROUPS", "0"))
    ncols = ksize * ksize * ntc
    partial = torch.empty((n_cig * n_cot, n_part, 128, ncols), dtype=torch.float32, device=x.t.device)
    a = _lib.WgradArgs()
    a.x, a.dy, a.partial = x.t.data_ptr(), dy.data_ptr(), partial.data_ptr()
    a.n_img, a.Z, a.Y, a.X, a.ksize = x.n_img, x.Z, x.Y, x.X, ksize
    a.TX, a.TY, a.TZ = TX, TY, TZ
    a.cig_blocks, a.cot_blocks, a.n_cig, a.n_cot = cig // 8, ntc // 8, n_cig, n_cot
    a.x_cbt, a.y_cbt, a.y_cb0, a.n_part = x.cbt, dy_cbt, dy_cb0, n_part
    for i, g in enumerate(groups):
        a.x_cb[i] = g
    if PROFILE is not None:
        _INFO[0] = {"flops": 2.0 * x.n_img * x.nvox * cin * cout_gemm * ksize ** 3,
                    "layer": f"wgrad k{ksize} cin{cin} n{cout_gemm} {x.Z}x{x.Y}x{x.X} img{x.n_img}",
                    "tile": (TX, TY, TZ, ntc, cig), "ctas": n_part * n_cig * n_cot}
    _call("mmseg_conv3d_wgrad", C.byref(a), _stream())
    dw = torch.empty(tuple(weight_shape), dtype=torch.float32, device=x.t.device)
    key = (segs_key, cig, x.t.device)
    cm = _CI_MAPS.get(key)
    if cm is None:  # cached: a host->device copy per call would also break CUDA-graph capture of the training step
        inv = [-1] * (n_cig * cig)            # accumulator row position -> weight input channel (-1: padding row)
        for ci_, pos in enumerate(ci_map):
            inv[pos] = ci_
        cm = _CI_MAPS[key] = torch.tensor(inv, dtype=torch.int32, device=x.t.device)
    cout = weight_shape[1] if transposed else weight_shape[0]
    if PROFILE is not None:
        _INFO[0] = {"bytes": 4.0 * partial.numel() + 4.0 * dw.numel(),
                    "layer": f"reduce-k{ksize}-cin{cin}-n{cout_gemm}-part{n_part}-pairs{n_cig * n_cot}"}
    _call("mmseg_wgrad_reduce", _ptr(partial), n_part, ksize, cig // 8, ntc // 8, n_cig, n_cot, cin, cout_gemm, cout,
          1 if transposed else 0, _ptr(cm), _ptr(dw), _stream())
    return dw


# --------------------------------------------------------------------------------------------- backward (elementwise)
def instnorm_act_bwd(raw: Tensor, mean_rstd: Tensor, n_img: int, channels: int, Z: int, Y: int, X: int,
                     gA: Optional[Blocked], gA_c0: int, gA_scale: float, gP: Optional[Blocked], gP_c0: int,
                     dx: Tensor, slope: float = 0.0, chan_scale: Optional[Tensor] = None,
                     chan_bias: Optional[Tensor] = None, dx_cbt: Optional[int] = None, dx_cb_off: int = 0,
                     between=None) -> None:
    """dx (blocked bf16 [n_img, channels/8, Z, Y, X, 8]) = gradient of the raw conv output; see mmseg_norm_bwd_args.
    dx_cbt / dx_cb_off: dx is a channel-block range of a wider blocked buffer (dx_cbt blocks per image)."""
    a = _lib.NormBwdArgs()
    nvox = Z * Y * X
    n_chunks = max(1, min(64, (nvox + 8191) // 8192))
    rows = n_img * channels // 8
    if rows * n_chunks > 296 and os.environ.get("MMSEG_BWD_WAVES", "1") == "1":
        # the reduce kernel runs two 256-thread CTAs per SM: round its grid to whole waves of 296 CTAs (a 512-CTA grid — the
        # 128^3, C = 32, B = 2 layers — otherwise spends its last 0.27 waves at half occupancy)
        waves = max(1, round(rows * n_chunks / 296.0))
        n_chunks = max(1, min(256, (296 * waves) // rows))
    partial = torch.empty((n_img * channels // 8, n_chunks, 16), dtype=torch.float32, device=raw.device)
    a.x, a.mean_rstd, a.partial, a.dx = raw.data_ptr(), mean_rstd.data_ptr(), partial.data_ptr(), dx.data_ptr()
    a.gA = gA.t.data_ptr() if gA is not None else None
    a.gP = gP.t.data_ptr() if gP is not None else None
    a.chan_scale = chan_scale.data_ptr() if chan_scale is not None else None
    a.chan_bias = chan_bias.data_ptr() if chan_bias is not None else None
    a.n_img, a.cb, a.Z, a.Y, a.X = n_img, channels // 8, Z, Y, X
    if gA is not None:
        a.gA_cbt, a.gA_cb_off = gA.cbt, gA_c0 // 8
    if gP is not None:
        a.gP_cbt, a.gP_cb_off = gP.cbt, gP_c0 // 8
    a.dx_cbt, a.dx_cb_off, a.n_chunks = (channels // 8 if dx_cbt is None else dx_cbt), dx_cb_off, n_chunks
    a.gA_scale, a.slope = gA_scale, slope
    if PROFILE is not None:
        _INFO[0] = {"bytes": n_img * channels * nvox * 4.0, "layer": f"bwd-reduce-c{channels}-{Z}-pool{int(gP is not None)}"}
    _call("mmseg_instnorm_act_bwd_reduce", C.byref(a), _stream())
    m12 = None
    if between is not None:
        # affine / group / batch norms: the caller turns the per-(image, channel) sums [n_img, C, 2] = (sum g', sum g' y^) into
        # the apply kernel's two subtraction terms (and takes its parameter gradients from the same sums)
        sums = partial.view(n_img, channels // 8, n_chunks, 2, 8).sum(2).permute(0, 1, 3, 2).reshape(n_img, channels, 2)
        m12 = between(sums).contiguous()
        assert m12.dtype == torch.float32 and tuple(m12.shape) == (n_img, channels, 2)
        a.m12 = m12.data_ptr()
    if PROFILE is not None:
        _INFO[0] = {"bytes": n_img * channels * nvox * 6.0, "layer": f"bwd-apply-c{channels}-{Z}-pool{int(gP is not None)}"}
    _call("mmseg_instnorm_act_bwd_apply", C.byref(a), _stream())


def unshuffle_k2s2(src: Blocked, c0: int, channels: int, dst: Tensor) -> None:
    """src: high-res blocked gradient (channels [c0, c0+channels)) -> dst blocked [n_img, 8*channels/8, Z/2, Y/2, X/2, 8]."""
    _call("mmseg_unshuffle_k2s2", _ptr(src.t), src.n_img, src.cbt, c0 // 8, channels // 8, src.Z // 2, src.Y // 2,
          src.X // 2, _ptr(dst), _stream())


def confusion_hist(pred: Tensor, target: Tensor, num_classes: int, counts: Tensor) -> None:
    """counts [K, K] int64 (rows = target, columns = prediction) += histogram of (target, pred) pairs."""
    _lib.require_device()
    assert pred.is_cuda and target.is_cuda and counts.dtype == torch.int64 and counts.is_contiguous()
    assert pred.dtype in (torch.int64, torch.uint8)
    pred, target = pred.contiguous(), target.contiguous().long()
    _call("mmseg_confusion_hist", _ptr(pred), 1 if pred.dtype == torch.uint8 else 0, _ptr(target), pred.numel(),
          num_classes, _ptr(counts), _stream())


# --------------------------------------------------------------------------------------------- token cross attention
def cross_attention(q: Blocked, q_c0: int, kv: Blocked, k_c0: int, v_c0: int, out: Blocked, o_c0: int, heads: int,
                    head_dim: int, scale: float, lse: Optional[Tensor] = None) -> None:
    """out[:, head h] = softmax(Q_h K_h^T * scale) V_h over all voxels; head_dim is the (padded) per-head channel count.
    lse (optional fp32 [n_img, heads, n_tok]): log-sum-exp rows, saved for cross_attention_bwd."""
    _lib.require_device()
    assert not (q.split or kv.split or out.split) and q.fmt == kv.fmt == out.fmt == _lib.FMT_BF16, \
        "the attention kernel runs in bf16 mode"
    assert q.nvox == kv.nvox == out.nvox and q.n_img == kv.n_img == out.n_img
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == q.n_img * heads * q.nvox
    if PROFILE is not None:
        _INFO[0] = {"flops": 4.0 * q.n_img * heads * q.nvox * q.nvox * head_dim,
                    "layer": f"attn h{heads} hd{head_dim} N{q.nvox} img{q.n_img}"}
    _call("mmseg_cross_attention_fwd", _ptr(q.t), q.cbt, q_c0 // 8, _ptr(kv.t), kv.cbt, k_c0 // 8, v_c0 // 8, _ptr(out.t),
          out.cbt, o_c0 // 8, q.n_img, heads, head_dim, q.nvox, scale, _ptr(lse), _stream())


def cross_attention_bwd(q: Blocked, q_c0: int, kv: Blocked, k_c0: int, v_c0: int, out: Blocked, o_c0: int, d_out: Blocked,
                        do_c0: int, lse: Tensor, dq: Blocked, dq_c0: int, dkv: Blocked, dk_c0: int, dv_c0: int, heads: int,
                        head_dim: int, scale: float) -> None:
    """dQ, dK, dV of cross_attention (blocked bf16, same head layout), by recomputation from q / kv / out / lse."""
    _lib.require_device()
    for b in (q, kv, out, d_out, dq, dkv):
        assert not b.split and b.fmt == _lib.FMT_BF16 and b.nvox == q.nvox and b.n_img == q.n_img
    dsum = torch.empty((q.n_img, heads, q.nvox), dtype=torch.float32, device=q.t.device)
    if PROFILE is not None:
        _INFO[0] = {"flops": 14.0 * q.n_img * heads * q.nvox * q.nvox * head_dim,
                    "layer": f"attn-bwd h{heads} hd{head_dim} N{q.nvox} img{q.n_img}"}
    _call("mmseg_cross_attention_bwd", _ptr(q.t), q.cbt, q_c0 // 8, _ptr(kv.t), kv.cbt, k_c0 // 8, v_c0 // 8, _ptr(out.t),
          out.cbt, o_c0 // 8, _ptr(d_out.t), d_out.cbt, do_c0 // 8, _ptr(lse), _ptr(dsum), _ptr(dq.t), dq.cbt, dq_c0 // 8,
          _ptr(dkv.t), dkv.cbt, dk_c0 // 8, dv_c0 // 8, q.n_img, heads, head_dim, q.nvox, scale, _stream())


def add_stats(a: Blocked, a_c0: int, b: Blocked, b_c0: int, channels: int):
    """(y fp32 blocked [n, channels/8, Z, Y, X, 8] = a + b, stats partial, n_chunks) for the residual + InstanceNorm."""
    n_chunks = max(1, min(64, (a.nvox + 4095) // 4096))
    y = torch.empty((a.n_img, channels // 8, a.Z, a.Y, a.X, 8), dtype=torch.float32, device=a.t.device)
    partial = torch.empty((a.n_img, n_chunks, channels, 2), dtype=torch.float32, device=a.t.device)
    _call("mmseg_add_stats", _ptr(a.t), a.cbt, a_c0 // 8, _ptr(b.t), b.cbt, b_c0 // 8, a.n_img, channels // 8, a.nvox,
          _ptr(y), _ptr(partial), n_chunks, _stream())
    return y, partial, n_chunks


def trilinear_resize(x: Tensor, size: Sequence[int]) -> Tensor:
    """F.interpolate(x, size, mode='trilinear', align_corners=True) for NCDHW fp32 (DeepSupervisionHead)."""
    assert x.dim() == 5 and x.dtype == torch.float32 and x.is_cuda
    x = x.contiguous()
    n, c, Zi, Yi, Xi = x.shape
    Zo, Yo, Xo = (int(v) for v in size)
    out = torch.empty((n, c, Zo, Yo, Xo), dtype=torch.float32, device=x.device)
    _call("mmseg_trilinear_resize", _ptr(x), n * c, Zi, Yi, Xi, _ptr(out), Zo, Yo, Xo, _stream())
    return out


def conv1x1_logits(src: Blocked, c0: int, cin: int, weight: Tensor, bias: Optional[Tensor], out: Tensor) -> None:
    """nn.Conv3d(cin, classes, 1) -> fp32 NCDHW logits on the CUDA cores (HBM-bound; fp32 weights, hi + lo inputs in
    parity mode).  weight: the module's [classes, cin, 1, 1, 1] parameter (read in place, no packed copy)."""
    cout = weight.shape[0]
    assert c0 % 8 == 0 and cin % 8 == 0 and cout <= 16 and weight.shape[1] == cin
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == src.n_img and out.shape[1] == cout
    w = weight.detach().reshape(cout, cin)
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    b = None
    if bias is not None:
        b = bias.detach()
        if b.dtype != torch.float32:
            b = b.float()
    if PROFILE is not None:
        _INFO[0] = {"flops": 2.0 * src.n_img * src.nvox * cin * cout, "bytes": src.n_img * src.nvox * (cin * 2.0 + cout * 4.0),
                    "layer": f"logits1x1 c{cin}->{cout} {src.Z}x{src.Y}x{src.X} img{src.n_img}"}
    _call("mmseg_conv1x1_logits", _ptr(src.t), src.n_img, src.cbt, c0 // 8, src.lo_off if src.split else 0, cin, src.nvox,
          _ptr(w), _ptr(b) if b is not None else None, cout, _ptr(out), src.fmt, _stream())


def modality_max(src: Blocked, M: int, channels: int, dst: Blocked, dst_c0: int = 0) -> None:
    assert not src.split and not dst.split and src.fmt == dst.fmt == _lib.FMT_BF16
    _call("mmseg_modality_max", _ptr(src.t), src.n_img, src.cbt, M, channels // 8, src.nvox, _ptr(dst.t), dst.cbt,
          dst_c0 // 8, _stream())


def modality_dot(stack: Blocked, M: int, channels: int, g: Blocked, g_c0: int) -> Tensor:
    """[n_img, M] fp32: sum over channels and voxels of g * stack[:, m] (gradient of the modality-gate weights)."""
    cb = channels // 8
    n_chunks = max(1, min(32, (stack.nvox + 8191) // 8192))
    partial = torch.empty((stack.n_img * M * cb, n_chunks), dtype=torch.float32, device=stack.t.device)
    out = torch.empty((stack.n_img, M), dtype=torch.float32, device=stack.t.device)
    _call("mmseg_modality_dot", _ptr(stack.t), stack.cbt, _ptr(g.t), g.cbt, g_c0 // 8, stack.n_img, M, cb, stack.nvox,
          _ptr(partial), n_chunks, _ptr(out), _stream())
    return out


# --------------------------------------------------------------------------------------------- Tversky / Focal losses
def tversky_fwd(logits: Tensor, target: Tensor, alpha: float, beta: float, smooth: float):
    _lib.require_device()
    B, Cc = logits.shape[:2]
    N = logits[0, 0].numel()
    n_blocks = max(1, min(148 * 4, (N + 255) // 256))
    partial = torch.empty((B, n_blocks, 3 * Cc + 2), dtype=torch.float32, device=logits.device)
    result = torch.empty(1, dtype=torch.float32, device=logits.device)
    sums = torch.empty((B, 3 * Cc + 2), dtype=torch.float32, device=logits.device)
    _call("mmseg_tversky_fwd", _ptr(logits), _ptr(target), B, Cc, N, alpha, beta, smooth, _ptr(partial), n_blocks,
          _ptr(result), _ptr(sums), _stream())
    return result, sums


def tversky_bwd(logits: Tensor, target: Tensor, sums: Tensor, grad_out: Tensor, alpha: float, beta: float, smooth: float) -> Tensor:
    B, Cc = logits.shape[:2]
    dl = torch.empty_like(logits)
    go = grad_out.reshape(1).float().contiguous()
    _call("mmseg_tversky_bwd", _ptr(logits), _ptr(target), B, Cc, logits[0, 0].numel(), alpha, beta, smooth, _ptr(sums),
          _ptr(go), _ptr(dl), _stream())
    return dl


def focal(logits: Tensor, target: Tensor, class_weights: Optional[Tensor], gamma: float,
          grad_out: Optional[Tensor] = None, backward: bool = False) -> Tensor:
    """forward: returns the scalar loss tensor [1]; backward=True: returns dlogits."""
    _lib.require_device()
    B, Cc = logits.shape[:2]
    N = logits[0, 0].numel()
    n_blocks = max(1, min(148 * 8, (B * N + 255) // 256))
    if backward:
        dl = torch.empty_like(logits)
        go = grad_out.reshape(1).float().contiguous()
        _call("mmseg_focal", _ptr(logits), _ptr(target), B, Cc, N, _ptr(class_weights), gamma, None, n_blocks, None,
              _ptr(go), _ptr(dl), _stream())
        return dl
    partial = torch.empty(n_blocks, dtype=torch.float32, device=logits.device)
    result = torch.empty(1, dtype=torch.float32, device=logits.device)
    _call("mmseg_focal", _ptr(logits), _ptr(target), B, Cc, N, _ptr(class_weights), gamma, _ptr(partial), n_blocks,
          _ptr(result), None, None, _stream())
    return result


# --------------------------------------------------------------------------------------------- SwinUNETR pieces (swin.cu)
def swin_patch_embed(x: Tensor, weight: Tensor, bias: Optional[Tensor], xs: Tensor) -> None:
    """Conv3d(Cin, F, k2, s2): NCDHW fp32 -> blocked fp32 tokens xs [n, F/8, Z, Y, X, 8]."""
    n, cin, Z2, Y2, X2 = x.shape
    F = weight.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and weight.dtype == torch.float32 and weight.is_contiguous()
    assert tuple(weight.shape[1:]) == (cin, 2, 2, 2) and not ((Z2 | Y2 | X2) & 1)
    assert xs.dtype == torch.float32 and xs.numel() == n * F * (Z2 // 2) * (Y2 // 2) * (X2 // 2)
    _call("mmseg_swin_patch_embed", _ptr(x), _ptr(weight), _ptr(bias), _ptr(xs), n, cin, F, Z2 // 2, Y2 // 2, X2 // 2, _stream())


def swin_layernorm(xs: Tensor, n_img: int, channels: int, voxels: int, dst: Optional[Blocked], dst_c0: int = 0,
                   add: Optional[Tensor] = None, gamma: Optional[Tensor] = None, beta: Optional[Tensor] = None,
                   eps: float = 1e-5) -> None:
    """xs += add; dst = LayerNorm over channels (optional affine).  xs / add: blocked fp32 [n, C/8, voxels, 8]."""
    assert xs.dtype == torch.float32 and xs.numel() == n_img * channels * voxels and channels % 8 == 0
    assert add is None or (add.dtype == torch.float32 and add.numel() >= xs.numel())
    fmt = dst.fmt if dst is not None else _lib.FMT_BF16
    if dst is not None:
        assert not dst.split and dst.nvox == voxels and dst.n_img == n_img and dst_c0 % 8 == 0
    _call("mmseg_swin_layernorm", _ptr(xs), _ptr(add), _ptr(gamma), _ptr(beta), _ptr(dst.t) if dst is not None else None,
          n_img, channels // 8, voxels, dst.cbt if dst is not None else 0, dst_c0 // 8, eps, fmt, _stream())


def swin_merge_ln(xs: Tensor, n_img: int, channels: int, Z: int, Y: int, X: int, gamma: Tensor, beta: Tensor, dst: Blocked,
                  eps: float = 1e-5) -> None:
    assert xs.dtype == torch.float32 and xs.numel() == n_img * channels * Z * Y * X
    assert dst.channels == 8 * channels and (dst.Z, dst.Y, dst.X) == (Z // 2, Y // 2, X // 2) and not dst.split
    _call("mmseg_swin_merge_ln", _ptr(xs), _ptr(gamma), _ptr(beta), _ptr(dst.t), n_img, channels // 8, Z, Y, X, eps, dst.fmt,
          _stream())


def swin_window_attention(qkv: Blocked, out: Blocked, table: Tensor, qkv_bias: Optional[Tensor], heads: int,
                          window: Sequence[int], shift: Sequence[int], out_c0: int = 0) -> None:
    """softmax(q k^T / 4 + relative-position bias [+ shift mask]) v per (window, head); qkv channels = [q | k | v]."""
    C_ = heads * 16
    assert qkv.channels == 3 * C_ and not qkv.split and not out.split and qkv.fmt == out.fmt
    assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[1] == heads
    a = _lib.SwinAttnArgs()
    a.qkv, a.out, a.table = qkv.t.data_ptr(), out.t.data_ptr(), table.data_ptr()
    a.qkv_bias = qkv_bias.data_ptr() if qkv_bias is not None else None
    a.n_img, a.D, a.H, a.W = qkv.n_img, qkv.Z, qkv.Y, qkv.X
    for i in range(3):
        a.window[i], a.shift[i] = int(window[i]), int(shift[i])
    a.heads, a.head_dim = heads, 16
    a.qkv_cbt, a.out_cbt, a.out_cb_off = qkv.cbt, out.cbt, out_c0 // 8
    a.scale, a.elem_fmt = 0.25, qkv.fmt
    if PROFILE is not None:
        _INFO[0] = {"flops": 4.0 * qkv.n_img * qkv.nvox * min(343, qkv.nvox) * C_,
                    "layer": f"winattn c{C_} {qkv.Z}x{qkv.Y}x{qkv.X} img{qkv.n_img} shift{int(shift[0])}"}
    _call("mmseg_swin_window_attention", C.byref(a), _stream())


def instnorm_residual_act(a: Tensor, a_is_f32: bool, a_mr: Tensor, r: Tensor, r_is_f32: bool, r_mr: Optional[Tensor],
                          r_cbt: int, r_c0: int, dst: Blocked, dst_c0: int, n_img: int, channels: int, voxels: int,
                          slope: float) -> None:
    """y = LeakyReLU(IN(a) + (IN(r) if r_mr is not None else r)) -> dst (UnetResBlock tail)."""
    assert not dst.split and dst.nvox == voxels
    if PROFILE is not None:
        nel = n_img * channels * voxels
        _INFO[0] = {"bytes": nel * ((4 if a_is_f32 else 2) + (4 if r_is_f32 else 2) + 2),
                    "layer": f"resnorm c{channels} vox{voxels} img{n_img}"}
    _call("mmseg_instnorm_residual_act", _ptr(a), 1 if a_is_f32 else 0, _ptr(a_mr), _ptr(r), 1 if r_is_f32 else 0,
          _ptr(r_mr), r_cbt, r_c0 // 8, _ptr(dst.t), dst.cbt, dst_c0 // 8, n_img, channels // 8, voxels, slope, dst.fmt,
          _stream())
