"""Training engine: forward with saved state + backward (dgrad / wgrad / norm-act-pool backward) for UNet3D and
DualEncoder, entirely in the sm_100a kernels (bf16 operands, fp32 accumulation, fp32 weight gradients).

The forward records a tape of ops; `backward(dlogits)` walks it in reverse.  Every activation buffer has a gradient
buffer of the same blocked shape; each region of a gradient buffer is written by exactly one dgrad launch, and the
two consumers of an encoder output (skip connection + MaxPool path) are merged inside the norm-backward kernel, so no
gradient accumulation pass (and no float atomic) exists anywhere — the whole step is deterministic
(works under torch.use_deterministic_algorithms(True), unlike the reference).

Reference semantics: autograd through UNet3D.forward (src/models/backbones/unet.py:165-200) / DualEncoder.forward
(src/models/backbones/dual_encoder.py:112-199) as driven by Trainer._train_epoch (src/trainer/trainer.py:237-243).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import os

import torch

from . import _lib
from . import kernels as K
from .kernels import Blocked

Tensor = torch.Tensor


def _wrap(t: Tensor, n: int, channels: int, Z: int, Y: int, X: int) -> Blocked:
    """A Blocked view over an existing bf16 tensor [n, channels/8, Z, Y, X, 8] (no allocation)."""
    b = Blocked.__new__(Blocked)
    b.n_img, b.channels, b.Z, b.Y, b.X = n, channels, Z, Y, X
    b.cb = b.cbt = channels // 8
    b.split, b.lo_off = False, 0
    b.fmt, b.nm = _lib.FMT_BF16, None
    b.t = t
    return b


class TrainEngine:
    def __init__(self, module, kind: str):
        assert kind in ("unet", "dual")
        self.module, self.kind = module, kind
        self._shape = None
        self.A: Dict[str, Blocked] = {}      # activations
        self.G: Dict[str, Blocked] = {}      # gradients w.r.t. activations
        self.S: Dict[str, Tensor] = {}       # per-op saved tensors (raw conv outputs, mean/rstd) and workspaces
        self.tape: List[dict] = []
        self.grads: Dict[Tensor, Tensor] = {}
        self._wstream = None
        self._reducer = None
        self._handed = set()
        self._buf_busy: Dict[str, object] = {}
        self._flip = 0
        self._plans: Dict[Tuple, K.PackPlan] = {}
        self._batch = None
        self._batch_keys: list = []
        self._fresh: Dict[Tuple, int] = {}   # plan key -> parameter version its packed operand was derived from

    def _packed(self, param: Tensor, variant: str, make) -> K.PackedConv:
        """Kernel-layout operand of `param` in form `variant`, re-derived from the live parameter by the step's ONE
        batched repack launch (kernels.PackBatch / mmseg_weights_repack_multi, issued by `_repack_all` at the start of
        the forward); a plan that did not exist yet at that point (first step) runs its own launch."""
        key = (id(param), variant)
        plan = self._plans.get(key)
        if plan is None or plan.weight.data_ptr() != param.data_ptr():
            plan = self._plans[key] = make()
            self._batch = None
            self._fresh.pop(key, None)
        if self._fresh.get(key) == plan.weight._version:     # (detached views share the parameter's version counter)
            return plan.pc
        self._fresh[key] = plan.weight._version
        return plan.run()

    def _repack_all(self) -> None:
        """Every packed operand the previous step used, from the current parameters, in one launch."""
        if not self._plans or os.environ.get("MMSEG_REPACK_BATCH", "1") != "1":
            return
        if self._batch is None or not self._batch.valid():
            live = {k: p for k, p in self._plans.items()}
            self._batch = K.PackBatch(list(live.values()))
            self._batch_keys = list(live.keys())
        self._batch.run()
        for k, p in zip(self._batch_keys, self._batch.plans):
            self._fresh[k] = p.weight._version

    # ---------------------------------------------------------------- buffers
    def _reset(self, n, Z, Y, X, device):
        if self._shape != (n, Z, Y, X, str(device)):
            self.A.clear(); self.G.clear(); self.S.clear()
            self._shape = (n, Z, Y, X, str(device))
        self.device = device
        self.tape = []
        self.grads = {}

    def act(self, name: str, n, channels, Z, Y, X) -> Blocked:
        b = self.A.get(name)
        if b is None:
            b = self.A[name] = Blocked(n, channels, Z, Y, X, False, self.device)
            b.name = name
        return b

    def grad_of(self, a: Blocked) -> Blocked:
        g = self.G.get(a.name)
        if g is None:
            g = self.G[a.name] = Blocked(a.n_img, a.channels, a.Z, a.Y, a.X, False, self.device)
            g.name = "g_" + a.name
        return g

    def saved(self, name: str, shape, dtype) -> Tensor:
        t = self.S.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self.S[name] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    # ---------------------------------------------------------------- forward ops
    def conv_norm_act(self, name: str, src: Blocked, segs, conv, dst: Blocked, dst_c0: int = 0,
                      pooled: Optional[Blocked] = None, slope: float = 0.0, gspec=None, need_dgrad: bool = True,
                      chan_scale: Optional[Tensor] = None, gate_ref=None, norm=None):
        """norm: the block's norm module (reference unet.py:29-41).  None / InstanceNorm3d(affine=False): the fused path
        below; GroupNorm / BatchNorm3d / Identity: `_conv_generic_norm_act`."""
        from .engine import norm_kind_train
        nk = norm_kind_train(norm)
        if nk != "instance":
            return self._conv_generic_norm_act(name, src, segs, conv, dst, dst_c0, pooled, slope, need_dgrad, norm, nk,
                                               gspec=gspec, chan_scale=chan_scale, gate_ref=gate_ref)
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        cout = conv.weight.shape[0]
        seg_ch = tuple(s[1] for s in segs)
        pw = self._packed(conv.weight, ("fwd", seg_ch), lambda: K.PackPlan.forward(conv.weight, None, False, seg_ch, use_bias=False))
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], False)
        tile = K.plan_conv_norm((X, Y, Z), n, pw, False, a_cb)
        raw = self.saved(name + ".raw", (n, cout // 8, Z, Y, X, 8), torch.bfloat16)
        stats = self.saved("ws.stats", (n * tile.tiles_per_img * cout * 2,), torch.float32)
        mr = self.saved(name + ".mr", (n, cout, 2), torch.float32)
        K.conv3d(src, pw, a_cb, raw, _lib.OUT_BLOCKED_BF16, stats=stats, dst_cbt=cout // 8, tile=tile)
        if chan_scale is None and tile.tiles_per_img <= 64:
            # statistics finalized in the apply kernel's prologue; the (mean, rstd) table the backward needs is written
            # by the first block of every (image, channel-block) row
            K.instnorm_act_apply(raw, False, None, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, 0,
                                 stats=stats, tiles_per_img=tile.tiles_per_img, mean_rstd_out=mr)
        else:
            K.instnorm_finalize(stats, n, tile.tiles_per_img, cout, Z * Y * X, mr)
            mr_apply = mr
            if chan_scale is not None:  # Dropout3d: relu(y^) * s == relu(y^ * s) for s >= 0 -> fold s into rstd
                mr_apply = self.saved(name + ".mr_drop", (n, cout, 2), torch.float32)
                mr_apply.copy_(mr)
                mr_apply[:, :, 1] *= chan_scale
            K.instnorm_act_apply(raw, False, mr_apply, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, 0)
        self.tape.append(dict(kind="cna", name=name, src=src, segs=list(segs), conv=conv, dst=dst, dst_c0=dst_c0,
                              pooled=pooled, slope=slope, raw=raw, mr=mr, gspec=gspec, need_dgrad=need_dgrad,
                              chan_scale=chan_scale, gate_ref=gate_ref))


    # ---------------------------------------------------------------- GroupNorm / BatchNorm3d / Identity blocks
    def _conv_generic_norm_act(self, name, src, segs, conv, dst, dst_c0, pooled, slope, need_dgrad, norm, nk,
                               gspec=None, chan_scale=None, gate_ref=None):
        """ConvBlock3D half with model.backbone.norm = group | batch | anything else (Identity) — reference unet.py:29-41,
        53-60.  The conv keeps its bias (only InstanceNorm cancels it); the statistics come from the conv epilogue's
        partials and are combined over the norm's reduction set on [n, C] tensors; the apply kernel gets gamma folded into
        rstd and beta as its shift.  BatchNorm3d in train mode uses batch statistics and updates its running buffers."""
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        nvox = Z * Y * X
        cout = conv.weight.shape[0]
        seg_ch = tuple(s_[1] for s_ in segs)
        pw = self._packed(conv.weight, ("fwdb", seg_ch), lambda: K.PackPlan.forward(conv.weight, conv.bias, False, seg_ch))
        a_cb = K.a_chunk_table(src, [s_[0] for s_ in segs], list(seg_ch), False)
        tile = K.plan_conv_norm((X, Y, Z), n, pw, False, a_cb)
        raw = self.saved(name + ".raw", (n, cout // 8, Z, Y, X, 8), torch.bfloat16)
        stats = self.saved("ws.stats", (n * tile.tiles_per_img * cout * 2,), torch.float32)
        K.conv3d(src, pw, a_cb, raw, _lib.OUT_BLOCKED_BF16, stats=stats, dst_cbt=cout // 8, tile=tile)
        dev = raw.device
        ones, zeros = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        gamma = norm.weight.detach().float() if getattr(norm, "weight", None) is not None else ones
        beta = norm.bias.detach().float() if getattr(norm, "bias", None) is not None else zeros
        stat_grad = True          # do the statistics depend on this batch (so that gradient flows through them)?
        if nk == "none":
            mean, rstd, stat_grad = zeros.expand(n, cout), ones.expand(n, cout), False
        else:
            st = stats[:n * tile.tiles_per_img * cout * 2].view(n, tile.tiles_per_img, cout, 2).double().sum(1)   # [n, C, 2]
            if nk == "group":
                G = norm.num_groups
                sg = st.view(n, G, cout // G, 2).sum(2) / float(nvox * (cout // G))
                mu = sg[..., 0]
                var = (sg[..., 1] - mu * mu).clamp_min(0.0)
                mean = mu.repeat_interleave(cout // G, 1).float()
                rstd = (1.0 / torch.sqrt(var + norm.eps)).repeat_interleave(cout // G, 1).float()
            elif norm.training or not norm.track_running_stats:      # BatchNorm3d with batch statistics
                sb = st.sum(0) / float(n * nvox)
                mu = sb[:, 0]
                var = (sb[:, 1] - mu * mu).clamp_min(0.0)
                if norm.track_running_stats:
                    cnt = float(n * nvox)
                    mom = norm.momentum if norm.momentum is not None else 1.0 / float(norm.num_batches_tracked.item() + 1)
                    norm.running_mean.mul_(1 - mom).add_(mu.to(norm.running_mean.dtype), alpha=mom)
                    norm.running_var.mul_(1 - mom).add_((var * cnt / max(cnt - 1.0, 1.0)).to(norm.running_var.dtype), alpha=mom)
                    norm.num_batches_tracked += 1
                mean = mu.float().expand(n, cout)
                rstd = (1.0 / torch.sqrt(var + norm.eps)).float().expand(n, cout)
            else:                                                    # BatchNorm3d in eval mode: running statistics
                mean = norm.running_mean.detach().float().expand(n, cout)
                rstd = (1.0 / torch.sqrt(norm.running_var.detach().float() + norm.eps)).expand(n, cout)
                stat_grad = False
        mean, rstd = mean.contiguous(), rstd.contiguous()
        scale_fwd, shift = rstd * gamma, beta.expand(n, cout)
        if chan_scale is not None:      # Dropout3d's per-(sample, channel) factor s >= 0: s * act(z) = act(s * z)
            scale_fwd, shift = scale_fwd * chan_scale, shift * chan_scale
        mr_fwd = torch.stack([mean, scale_fwd], -1).contiguous()
        K.instnorm_act_apply(raw, False, mr_fwd, n, cout, Z, Y, X, dst, dst_c0, slope, pooled, 0, shift=shift.contiguous())
        self.tape.append(dict(kind="cna", name=name, src=src, segs=list(segs), conv=conv, dst=dst, dst_c0=dst_c0,
                              pooled=pooled, slope=slope, raw=raw, mr=None, gspec=gspec, need_dgrad=need_dgrad,
                              chan_scale=chan_scale, gate_ref=gate_ref, norm=norm, nk=nk, mean=mean, rstd=rstd, gamma=gamma,
                              beta=beta, stat_grad=stat_grad))

    def _bwd_generic_norm(self, op, draw_t, gA, gA_c0, gP, scale=1.0, chan_scale=None, chan_bias=None):
        """Backward of the norms above through the SAME two kernels as InstanceNorm: with (mean', rstd') = (mean - beta /
        (rstd gamma), rstd gamma) the kernels' normalised value IS the pre-activation z = y^ gamma + beta, so their sums are
        (sum h, sum h z) with h = g * act'(z); from those: dbeta = sum h, dgamma = sum h y^, and — combined over the norm's
        reduction set R (the group's channels / all images / nothing) —
            dx = rstd (h gamma - M1 - y^ M2),  M1 = mean_R(h gamma), M2 = mean_R(h gamma y^)
               = rstd' (h - A - z B),          A = M1 / gamma - beta M2 / gamma^2,  B = M2 / gamma^2,
        which is the apply kernel's formula with (A, B) as its two subtraction terms."""
        norm, nk = op["norm"], op["nk"]
        mean, rstd, gamma, beta = op["mean"], op["rstd"], op["gamma"], op["beta"]
        src, conv = op["src"], op["conv"]
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        nvox = float(Z * Y * X)
        cout = conv.weight.shape[0]
        g_safe = torch.where(gamma.abs() < 1e-12, torch.full_like(gamma, 1e-12), gamma)
        mr_b = torch.stack([mean - beta / (rstd * g_safe), rstd * g_safe], -1).contiguous()

        def between(sums):                     # sums [n, C, 2] = (sum h, sum h z)
            s_h, s_hz = sums[..., 0], sums[..., 1]
            s_hy = (s_hz - beta * s_h) / g_safe
            if getattr(norm, "weight", None) is not None:
                self.grads[norm.weight] = s_hy.sum(0)
            if getattr(norm, "bias", None) is not None:
                self.grads[norm.bias] = s_h.sum(0)
            if not op["stat_grad"]:
                return torch.zeros_like(sums)
            t1, t2 = s_h * gamma, s_hy * gamma
            if nk == "group":
                G = norm.num_groups
                cnt = nvox * (cout // G)
                M1 = (t1.view(n, G, -1).sum(2) / cnt).repeat_interleave(cout // G, 1)
                M2 = (t2.view(n, G, -1).sum(2) / cnt).repeat_interleave(cout // G, 1)
            else:                              # batch statistics: over images and voxels
                M1 = (t1.sum(0) / (n * nvox)).expand(n, cout)
                M2 = (t2.sum(0) / (n * nvox)).expand(n, cout)
            return torch.stack([M1 / g_safe - beta * M2 / (g_safe * g_safe), M2 / (g_safe * g_safe)], -1)

        # h = (scale * chan_scale * gA [+ routed gP] + chan_bias) * act'(z): the fused gradient's share of this modality
        # (mean / add: `scale`; attention gate: its weight and the pooled-mean term), Dropout3d's factor
        K.instnorm_act_bwd(op["raw"], mr_b, n, cout, Z, Y, X, gA, gA_c0, scale, gP, 0, draw_t, op["slope"],
                           chan_scale, chan_bias, between=between)

    def conv_transpose(self, name: str, src: Blocked, up, dst: Blocked):
        cin = up.weight.shape[0]
        pw = self._packed(up.weight, "convt", lambda: K.PackPlan.forward(up.weight, up.bias, False, None, transposed=True))
        a_cb = K.a_chunk_table(src, [0], [cin], False)
        K.conv3d(src, pw, a_cb, dst.t, _lib.OUT_CONVT_K2S2, dst_cbt=dst.cbt, dst_cb_off=0, dst_lo_off=0)
        self.tape.append(dict(kind="convt", name=name, src=src, up=up, dst=dst))

    def conv_logits(self, name: str, src: Blocked, conv, logits: Tensor):
        cin = conv.weight.shape[1]
        if cin % 8 == 0 and cin <= 256 and conv.out_channels <= 16:
            K.conv1x1_logits(src, 0, cin, conv.weight, conv.bias, logits)
        else:
            pw = K.pack_conv_weight(conv.weight, conv.bias, False, None)
            a_cb = K.a_chunk_table(src, [0], [cin], False)
            K.conv3d(src, pw, a_cb, logits, _lib.OUT_NCDHW_F32)
        self.tape.append(dict(kind="logits", name=name, src=src, conv=conv))

    def conv_bias(self, name: str, src: Blocked, segs, conv, dst: Blocked, dst_c0: int):
        """1x1 conv + bias straight into an activation buffer (DualEncoder 'concat' fusion_proj)."""
        seg_ch = tuple(s[1] for s in segs)
        pw = self._packed(conv.weight, ("fwdb", seg_ch), lambda: K.PackPlan.forward(conv.weight, conv.bias, False, seg_ch))
        a_cb = K.a_chunk_table(src, [s[0] for s in segs], [s[1] for s in segs], False)
        K.conv3d(src, pw, a_cb, dst.t, _lib.OUT_BLOCKED_BF16, dst_cbt=dst.cbt, dst_cb_off=dst_c0 // 8)
        self.tape.append(dict(kind="convb", name=name, src=src, segs=list(segs), conv=conv, dst=dst, dst_c0=dst_c0))

    def gate_fuse(self, name: str, stack: Blocked, M: int, C: int, att, dst: Blocked, dst_c0: int) -> dict:
        """CrossModalAttention forward (dual_encoder.py:243-254): channel means -> MLP + softmax -> weighted sum."""
        pooled = K.channel_mean(stack, 0, M * C)
        w = K.gate_mlp(pooled, att[2].weight, att[2].bias, att[4].weight, att[4].bias)
        K.modality_combine(stack, M, C, dst, dst_c0, w)
        op = dict(kind="gate", name=name, stack=stack, M=M, C=C, att=att, dst=dst, dst_c0=dst_c0, pooled=pooled, w=w)
        self.tape.append(op)
        return op

    def _bwd_gate(self, op):
        stack, M, C, att = op["stack"], op["M"], op["C"], op["att"]
        g = self.grad_of(op["dst"])
        dw = K.modality_dot(stack, M, C, g, op["dst_c0"])                     # [n, M]
        # backward of the gate MLP ([n, M*C] -> hidden -> [n, M] -> softmax) in two kernel launches
        params = [att[2].weight, att[2].bias, att[4].weight, att[4].bias]
        dpooled, dw1, db1, dw2, db2 = K.gate_mlp_bwd(op["pooled"], *params, dw)
        op["dpooled"] = dpooled
        for p, gp in zip(params, (dw1, db1, dw2, db2)):
            self.grads[p] = gp

    # ---------------------------------------------------------------- backward ops
    def _dgrad(self, dy: Blocked, dy_channels: int, pw: K.PackedConv, dst: Blocked, dst_c0: int, dy_c0: int = 0):
        """dst[:, dst_c0 : dst_c0 + Cout'] = conv(dy[:, dy_c0 : dy_c0 + dy_channels], pw) with the forward tcgen05 kernel
        (pw = the dgrad-form operand: flipped taps, channels transposed)."""
        a_cb = K.a_chunk_table(dy, [dy_c0], [dy_channels], False)
        K.conv3d(dy, pw, a_cb, dst.t, _lib.OUT_BLOCKED_BF16, dst_cbt=dst.cbt, dst_cb_off=dst_c0 // 8)

    def _wgrad_async(self, param, buf_key: str, *wargs, **wkw) -> None:
        """Weight gradient on a SIDE stream: it only depends on the raw-output gradient just produced, so it overlaps
        the dgrad / norm-backward kernels of the layers below (tensor-bound wgrad next to HBM-bound norm backward).
        `buf_key` names the gradient workspace it reads; the next writer of that workspace waits for this wgrad."""
        main = torch.cuda.current_stream(self.device)
        if os.environ.get("MMSEG_WGRAD_SIDE_STREAM", "1") == "0":   # profiling: serialise, so per-kernel times are clean
            g = K.conv3d_wgrad(*wargs, **wkw)
            self.grads[param] = g
            if self._reducer is not None and param.requires_grad:
                self._reducer.grad_ready(param, g)
                self._handed.add(param)
            return
        if self._wstream is None:
            self._wstream = torch.cuda.Stream(device=self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self._wstream.wait_event(ready)
        with torch.cuda.stream(self._wstream):
            g = K.conv3d_wgrad(*wargs, **wkw)
            g.record_stream(main)
            self.grads[param] = g
            if self._reducer is not None and param.requires_grad:
                self._reducer.grad_ready(param, g)       # packs + maybe launches the bucket all-reduce behind the wgrad
                self._handed.add(param)
            done = torch.cuda.Event()
            done.record(self._wstream)
        self._buf_busy[buf_key] = done

    def _wait_buf(self, buf_key: str) -> None:
        ev = self._buf_busy.pop(buf_key, None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _channel_sums(self, b: Blocked, c0: int, channels: int) -> Tensor:
        """[channels] fp32 sums over images and voxels (bias gradients)."""
        return (K.channel_mean(b, c0, channels) * float(b.nvox)).sum(0)

    def _bwd_cna(self, op):
        src, dst, conv = op["src"], op["dst"], op["conv"]
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        if op["gspec"] is not None:
            gA, gA_c0, scale = op["gspec"]
        else:
            gA, gA_c0, scale = self.grad_of(dst), op["dst_c0"], 1.0
        gP = self.grad_of(op["pooled"]) if op["pooled"] is not None else None
        self._flip ^= 1
        dkey = f"ws.draw{self._flip}"            # two workspaces: the side-stream wgrad of the previous layer may still
        self._wait_buf(dkey)                     # be reading the other one
        draw_t = self.saved(dkey, (n * cout * Z * Y * X,), torch.bfloat16).view(n, cout // 8, Z, Y, X, 8)
        chan_scale, chan_bias = op["chan_scale"], None
        if op.get("gate_ref") is not None:   # encoder output feeding a CrossModalAttention gate (modality i)
            gate, i = op["gate_ref"]
            chan_scale = gate["w"][:, i:i + 1].expand(n, cout).contiguous()
            chan_bias = (gate["dpooled"][:, i * cout:(i + 1) * cout] / float(Z * Y * X)).contiguous()
        generic = op.get("nk", "instance") != "instance"
        if generic:
            self._bwd_generic_norm(op, draw_t, gA, gA_c0, gP, scale, chan_scale, chan_bias)
        else:
            K.instnorm_act_bwd(op["raw"], op["mr"], n, cout, Z, Y, X, gA, gA_c0, scale, gP, 0, draw_t, op["slope"],
                               chan_scale, chan_bias)
        draw = _wrap(draw_t, n, cout, Z, Y, X)
        ks = conv.weight.shape[2]
        self._wgrad_async(conv.weight, dkey, src, op["segs"], draw_t, cout // 8, 0, cout, ks, conv.weight.shape)
        if conv.bias is not None:
            if generic:   # the conv bias is live under every norm but InstanceNorm
                self.grads[conv.bias] = self._channel_sums(draw, 0, cout)
            else:         # cancelled exactly by the InstanceNorm mean subtraction
                self.grads[conv.bias] = torch.zeros_like(conv.bias, dtype=torch.float32)
        if op["need_dgrad"]:
            segs = op["segs"]
            # (a single segment may have any channel count: the first layers' 1- or 2-channel input; the padded output
            # channels of the dgrad conv are exact zeros)
            assert (len(segs) == 1 and segs[0][0] % 8 == 0) or \
                (all(s[1] % 16 == 0 for s in segs) and all(segs[i][0] + segs[i][1] == segs[i + 1][0] for i in range(len(segs) - 1))), \
                "dgrad needs contiguous 16-channel-aligned input segments"
            pwd = self._packed(conv.weight, "dgrad", lambda: K.PackPlan.dgrad(conv.weight))   # flipped, [cin, cout, k, k, k]
            self._dgrad(draw, cout, pwd, self.grad_of(src), segs[0][0])

    def _bwd_convt(self, op):
        src, dst, up = op["src"], op["dst"], op["up"]
        cin, f = up.weight.shape[0], up.weight.shape[1]
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        gdst = self.grad_of(dst)
        self._wait_buf("ws.dyu")
        dyu_t = self.saved("ws.dyu", (n * 8 * f * Z * Y * X,), torch.bfloat16).view(n, f, Z, Y, X, 8)  # 8f/8 = f blocks
        K.unshuffle_k2s2(gdst, 0, f, dyu_t)
        dyu = _wrap(dyu_t, n, 8 * f, Z, Y, X)
        self._wgrad_async(up.weight, "ws.dyu", src, [(0, cin)], dyu_t, f, 0, 8 * f, 1, up.weight.shape, transposed=True)
        if up.bias is not None:
            self.grads[up.bias] = self._channel_sums(dyu, 0, 8 * f).view(8, f).sum(0)
        pwd = self._packed(up.weight, "convt_dgrad", lambda: K.PackPlan.convt_dgrad(up.weight))
        self._dgrad(dyu, 8 * f, pwd, self.grad_of(src), 0)

    def _bwd_logits(self, op, dlogits: Tensor):
        src, conv = op["src"], op["conv"]
        k, cin = conv.weight.shape[0], conv.weight.shape[1]
        n, Z, Y, X = src.n_img, src.Z, src.Y, src.X
        kp = (k + 15) // 16 * 16
        dl = self.act("ws.dlogits", n, kp, Z, Y, X)
        K.pack_ncdhw(dlogits.contiguous().float(), dl)
        self.grads[conv.weight] = K.conv3d_wgrad(src, [(0, cin)], dl.t, dl.cbt, 0, k, 1, conv.weight.shape)
        if conv.bias is not None:
            self.grads[conv.bias] = self._channel_sums(dl, 0, kp)[:k]
        pwd = self._packed(conv.weight, "k1_dgrad", lambda: K.PackPlan.k1_dgrad(conv.weight))
        self._dgrad(dl, k, pwd, self.grad_of(src), 0)

    def _bwd_convb(self, op):
        src, dst, conv, c0 = op["src"], op["dst"], op["conv"], op["dst_c0"]
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        g = self.grad_of(dst)
        self.grads[conv.weight] = K.conv3d_wgrad(src, op["segs"], g.t, g.cbt, c0 // 8, cout, 1, conv.weight.shape)
        if conv.bias is not None:
            self.grads[conv.bias] = self._channel_sums(g, c0, cout)
        # dgrad reads the gradient region in place: view it as a blocked tensor through the K-chunk table
        pwd = self._packed(conv.weight, "k1_dgrad", lambda: K.PackPlan.k1_dgrad(conv.weight))
        self._dgrad(g, cout, pwd, self.grad_of(src), op["segs"][0][0], dy_c0=c0)

    @torch.no_grad()
    def backward(self, dlogits: Tensor, reducer=None) -> Dict[Tensor, Tensor]:
        self._reducer = reducer
        handed = self._handed = set()
        self._buf_busy = {}
        for op in reversed(self.tape):
            if reducer is not None:  # gradients produced by the previous op go out while this op's kernels are queued
                for p, g in self.grads.items():
                    if p not in handed and p.requires_grad:
                        reducer.grad_ready(p, g)
                        handed.add(p)
            kind = op["kind"]
            if kind == "cna":
                self._bwd_cna(op)
            elif kind == "convt":
                self._bwd_convt(op)
            elif kind == "logits":
                self._bwd_logits(op, dlogits)
            elif kind == "convb":
                self._bwd_convb(op)
            elif kind == "gate":
                self._bwd_gate(op)
        if self._wstream is not None:   # join the side stream: every weight gradient is complete on the caller's stream
            torch.cuda.current_stream(self.device).wait_stream(self._wstream)
        if reducer is not None:
            for p, g in self.grads.items():
                if p not in handed and p.requires_grad:
                    reducer.grad_ready(p, g)
        self._reducer = None
        self.x_grad = None
        if getattr(self, "_input_grad", False):   # gradient w.r.t. the input volume, NCDHW fp32, modalities in channel order
            parts = [self.grad_of(a).to_ncdhw(0, c) for a, c in self._inputs]
            self.x_grad = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        return self.grads

    # ---------------------------------------------------------------- model forwards
    @torch.no_grad()
    def forward(self, x: Tensor, drop_scale: Optional[Tensor] = None, input_grad: bool = False) -> Tensor:
        """x NCDHW fp32 (CUDA) -> logits NCDHW fp32; records the tape.  drop_scale [n, f0]: Dropout3d mask * 1/(1-p).
        input_grad: also run the first layers' dgrad in backward() (the gradient w.r.t. x: autograd's input gradient)."""
        self._input_grad = bool(input_grad)
        self._inputs: List[Tuple[Blocked, int]] = []      # (blocked input buffer, real channels) in channel order
        _lib.require_device()
        if not x.is_cuda:
            raise RuntimeError("mmseg_b200 engines run on CUDA tensors only (no CPU fallback)")
        x = x.contiguous().float()
        n, _, Z, Y, X = x.shape
        self._reset(n, Z, Y, X, x.device)
        self._repack_all()
        m = self.module
        f, L = m.features, len(m.features)
        if any(d % (1 << (L - 1)) for d in (Z, Y, X)):
            raise NotImplementedError(f"spatial size {(Z, Y, X)} is not divisible by {1 << (L - 1)}")
        A = lambda name, ch, l: self.act(name, n, ch, Z >> l, Y >> l, X >> l)
        for l in range(L - 1):
            A(f"cat{l}", 2 * f[l], l)
        bott = A("bott", f[L - 1], L - 1)
        fused = lambda l: (bott, 0) if l == L - 1 else (self.A[f"cat{l}"], f[l])

        if self.kind == "unet":
            cin = m.in_channels
            a_in = A("in", (cin + 15) // 16 * 16, 0)
            K.pack_ncdhw(x, a_in)
            self._inputs.append((a_in, cin))
            self.conv_norm_act("init.c1", a_in, [(0, cin)], m.init_conv.conv1, A("e0.mid", f[0], 0),
                               need_dgrad=self._input_grad, norm=m.init_conv.norm1)
            for l in range(L):
                blk = m.init_conv if l == 0 else m.encoders[l - 1].conv
                if l > 0:
                    self.conv_norm_act(f"e{l}.c1", self.A[f"pool{l}"], [(0, f[l - 1])], blk.conv1, A(f"e{l}.mid", f[l], l),
                                       norm=blk.norm1)
                dst, c0 = fused(l)
                self.conv_norm_act(f"e{l}.c2", self.A[f"e{l}.mid"], [(0, f[l])], blk.conv2, dst, c0,
                                   pooled=A(f"pool{l + 1}", f[l], l + 1) if l < L - 1 else None, norm=blk.norm2)
            decoders = m.decoders
        else:
            M, cpm = m.num_modalities, m.in_channels_per_modality
            for l in range(L):
                A(f"stack{l}", M * f[l], l)
            scale = {"add": 1.0}.get(m.fusion_type, 1.0 / M)
            enc_last: Dict[int, list] = {}
            for i in range(M):
                a_in = A(f"m{i}.in", (cpm + 15) // 16 * 16, 0)
                K.pack_ncdhw(x[:, i * cpm:(i + 1) * cpm].contiguous(), a_in)
                self._inputs.append((a_in, cpm))
                enc = m.encoders[i]
                self.conv_norm_act(f"m{i}.init.c1", a_in, [(0, cpm)], enc["init_conv"].conv1, A(f"m{i}.e0.mid", f[0], 0),
                                   need_dgrad=self._input_grad, norm=enc["init_conv"].norm1)
                for l in range(L):
                    blk = enc["init_conv"] if l == 0 else enc["blocks"][l - 1].conv
                    if l > 0:
                        self.conv_norm_act(f"m{i}.e{l}.c1", self.A[f"m{i}.pool{l}"], [(0, f[l - 1])], blk.conv1,
                                           A(f"m{i}.e{l}.mid", f[l], l), norm=blk.norm1)
                    gspec = None
                    if m.fusion_type == "attention":   # scale / bias tables come from the gate op's backward
                        fd, fc0 = fused(l)
                        gspec = (self.grad_of(fd), fc0, 1.0)
                    elif m.fusion_type != "concat":  # mean / add: the stack gradient is the fused gradient times `scale`
                        fd, fc0 = fused(l)
                        gspec = (self.grad_of(fd), fc0, scale)
                    self.conv_norm_act(f"m{i}.e{l}.c2", self.A[f"m{i}.e{l}.mid"], [(0, f[l])], blk.conv2,
                                       self.A[f"stack{l}"], i * f[l],
                                       pooled=A(f"m{i}.pool{l + 1}", f[l], l + 1) if l < L - 1 else None, gspec=gspec,
                                       norm=blk.norm2)
                    if m.fusion_type == "attention":
                        enc_last.setdefault(l, []).append((self.tape[-1], i))
            for l in range(L):
                dst, c0 = fused(l)
                st = self.A[f"stack{l}"]
                if m.fusion_type == "concat":
                    self.conv_bias(f"fuse{l}", st, [(i * f[l], f[l]) for i in range(M)], m.fusion_proj[l], dst, c0)
                elif m.fusion_type == "attention":
                    gate = self.gate_fuse(f"gate{l}", st, M, f[l], m.fusion_layers[l].attention, dst, c0)
                    for op_i, i in enc_last[l]:
                        op_i["gate_ref"] = (gate, i)
                else:
                    K.modality_combine(st, M, f[l], dst, c0, None, scale)
            decoders = m.decoder

        cur = bott
        for j in range(L - 1):
            l = L - 2 - j
            dec = decoders[j]
            cat = self.A[f"cat{l}"]
            self.conv_transpose(f"d{l}.up", cur, dec.up, cat)
            self.conv_norm_act(f"d{l}.c1", cat, [(0, f[l]), (f[l], f[l])], dec.conv.conv1, A(f"d{l}.mid", f[l], l),
                               norm=dec.conv.norm1)
            last = l == 0
            self.conv_norm_act(f"d{l}.c2", self.A[f"d{l}.mid"], [(0, f[l])], dec.conv.conv2, A(f"d{l}.out", f[l], l),
                               chan_scale=drop_scale if last else None, norm=dec.conv.norm2)
            cur = self.A[f"d{l}.out"]
        logits = torch.empty((n, m.out_channels, Z, Y, X), dtype=torch.float32, device=x.device)
        self.conv_logits("out", cur, m.out_conv, logits)
        return logits


# set by Trainer.train_step on the stepping micro-batch of a data-parallel run (parallel.GradBucketReducer)
ACTIVE_REDUCER = None


class _ModelFunction(torch.autograd.Function):
    """logits = model(x) with the whole backward in the sm_100a kernels; parameters receive fp32 gradients."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, x: Tensor, drop_scale: Optional[Tensor], *params: Tensor) -> Tensor:
        ctx.engine, ctx.params = engine, params
        ctx.x_dtype = x.dtype
        return engine.forward(x, drop_scale, input_grad=x.requires_grad)

    @staticmethod
    def backward(ctx, dlogits: Tensor):
        red = ACTIVE_REDUCER if (ACTIVE_REDUCER is not None and ACTIVE_REDUCER.armed) else None
        grads = ctx.engine.backward(dlogits, red)
        out = []
        for p in ctx.params:
            g = grads.get(p)
            # armed data-parallel step: the reducer owns the gradient (bucket -> all-reduce -> .grad in finish())
            out.append(None if (g is None or not p.requires_grad or red is not None) else g.to(p.dtype).view_as(p))
        gx = ctx.engine.x_grad if ctx.needs_input_grad[1] else None
        return (None, None if gx is None else gx.to(ctx.x_dtype), None, *out)


def train_forward(engine: TrainEngine, x: Tensor, drop_scale: Optional[Tensor] = None) -> Tensor:
    params = [p for p in engine.module.parameters()]
    return _ModelFunction.apply(engine, x, drop_scale, *params)
