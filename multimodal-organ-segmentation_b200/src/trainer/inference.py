"""Sliding-window inference on B200 — drop-in for the call at reference src/trainer/trainer.py:381-392.

`sliding_window_inference` keeps MONAI's call signature (the only API the reference uses) and semantics
(SURVEY.md Appendix C): window grid with clamped last window, constant / gaussian importance map with the
max(min, 1e-3) floor, weighted accumulation in window order, division by the count map.

Engine path (predictor is a drop-in model with `.engine()`): windows are gathered straight into the engine's blocked
input buffer by a kernel, evaluated `engine_batch` at a time (the result does not depend on the batch size because
InstanceNorm statistics are per window), and blended by an owner-thread kernel in window order — no atomics, same
accumulation order as the reference loop.  One batch = gather + forward + blend is captured in a CUDA graph.

Multi-GPU (`shard_windows`): the ordered window list is cut into `world` contiguous chunks (axis-0 slabs); every
rank accumulates its chunk locally, then each rank receives — in rank order — the other ranks' partial sums that
fall into the axis-0 slab it owns, finalises that slab and the uint8 labels are all-gathered.
"""
import math
from typing import Callable, List, Optional, Sequence, Tuple

import contextlib
import gc
import os

import torch

from ... import kernels as K
from ... import _lib

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------ window arithmetic (host)

@contextlib.contextmanager
def capture_guard():
    """No cyclic garbage collection while a stream is capturing: collecting an unreachable inferer of an earlier call
    there would destroy ITS CUDA graphs (cudaGraphExecDestroy / cudaFree of the graph pool), which CUDA forbids during a
    capture and which invalidates it ("operation failed due to a previous error during capture").  torch.cuda.graph no
    longer runs gc.collect() itself, so collect first, then keep the collector off for the duration of the capture."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


def _scan_interval(image_size, roi_size, overlap) -> List[int]:
    out = []
    for im, r in zip(image_size, roi_size):
        if r == im:
            out.append(int(r))
        else:
            iv = int(r * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return out


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[Tuple[int, ...]]:
    """Window origins in MONAI order (axis 0 slowest, last axis fastest), last window clamped to the border."""
    per_axis = []
    for im, r, iv in zip(image_size, roi_size, _scan_interval(image_size, roi_size, overlap)):
        num = int(math.ceil(float(im) / iv))
        n = 1
        for d in range(num):
            if d * iv + r >= im:
                n = d + 1
                break
        per_axis.append([min(d * iv, im - r) for d in range(n)])
    out: List[Tuple[int, ...]] = [()]
    for starts in per_axis:
        out = [o + (s,) for o in out for s in starts]
    return out


def importance_tables(roi_size: Sequence[int], mode: str, sigma_scale: float = 0.125):
    """Separable factors of the importance map and the clamp floor.

    gaussian: w[z,y,x] = max((g_z*g_y)*g_x, floor), g[i] = exp(-(i-(r-1)/2)^2 / (2 (sigma_scale r)^2)) in fp32,
    floor = max(min(w), 1e-3) — identical arithmetic (and association order) to the fp32 tensor MONAI builds.
    """
    tabs = []
    if mode == "constant":
        for r in roi_size:
            tabs.append(torch.ones(r, dtype=torch.float32))
        return tabs, 1.0
    if mode != "gaussian":
        raise ValueError(f"unsupported blend mode {mode!r} (constant | gaussian)")
    for r in roi_size:
        sigma = r * sigma_scale
        x = torch.arange(-(r - 1) / 2.0, (r - 1) / 2.0 + 1, dtype=torch.float32)
        tabs.append(torch.exp(x ** 2 / (-2 * sigma ** 2)))
    wmin = (tabs[0].min() * tabs[1].min()) * tabs[2].min()
    return tabs, max(float(wmin), 1e-3)


def shard_windows(n_windows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of the ordered window list owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owned_slabs(starts: Sequence[Tuple[int, ...]], roi0: int, size0: int, world: int) -> List[Tuple[int, int]]:
    """Axis-0 partition [lo, hi) per rank: rank r owns from its first window's origin to the next rank's."""
    cuts = []
    for r in range(world):
        lo, hi = shard_windows(len(starts), world, r)
        cuts.append(min(s[0] for s in starts[lo:hi]) if hi > lo else None)
    out = []
    for r in range(world):
        if cuts[r] is None:
            out.append((size0, size0))
            continue
        lo = 0 if r == 0 else cuts[r]
        nxt = next((cuts[q] for q in range(r + 1, world) if cuts[q] is not None), size0)
        out.append((lo, max(lo, nxt)))
    # make the slabs a partition even when consecutive ranks start at the same origin
    fixed, prev_hi = [], 0
    for lo, hi in out:
        lo = max(lo, prev_hi) if hi > prev_hi else prev_hi
        hi = max(hi, lo)
        fixed.append((lo, hi))
        prev_hi = hi
    fixed[-1] = (fixed[-1][0], size0)
    return fixed


def touched_range(starts: Sequence[Tuple[int, ...]], lo: int, hi: int, roi0: int) -> Tuple[int, int]:
    """Axis-0 range written by windows [lo, hi)."""
    if hi <= lo:
        return (0, 0)
    s = [st[0] for st in starts[lo:hi]]
    return min(s), max(s) + roi0


def pick_engine_batch(n_windows: int, preferred: int = 12, lo: int = 6, hi: int = 16) -> int:
    """Windows per engine batch: the divisor of `n_windows` in [lo, hi] closest to `preferred` (every batch is then a
    full, graph-replayed batch — e.g. 75 windows per rank at 8 GPUs run as 5 x 15 instead of 9 x 8 + 3), else
    `preferred` (the shorter tail batch gets its own captured graph)."""
    if n_windows <= preferred:
        return max(1, n_windows)
    divs = [d for d in range(lo, hi + 1) if n_windows % d == 0]
    if not divs:
        return preferred
    return min(divs, key=lambda d: (abs(d - preferred), -d))


# ------------------------------------------------------------------------------------------ the inferer
class SlidingWindowInferer:
    """Stateful sliding-window engine of one model for one (roi, overlap, blend mode).

    Persistent across calls and volume shapes: the batch slots (engine, stream, window-origin slot, activation buffers —
    they depend on the roi, not on the volume).  Per volume shape: the window list and the accumulator.  CUDA graphs of
    a batch (gather + forward) are cached per (slot, batch size, volume buffer) and dropped whenever the model's
    parameters or norm buffers change (engine.module_version), so predict -> train / load_state_dict -> predict never
    replays a forward that reads stale packed weights."""

    def __init__(self, model, roi_size=(96, 96, 96), overlap: float = 0.5, mode: str = "constant",
                 sigma_scale: float = 0.125, engine_batch: Optional[int] = None, use_graph: bool = True):
        self.model = model
        self.roi = tuple(int(r) for r in roi_size)
        self.overlap, self.mode, self.sigma_scale = overlap, mode, sigma_scale
        self.engine_batch = engine_batch          # None: chosen per call from the number of windows (pick_engine_batch)
        self.use_graph = use_graph
        self.n_slots = int(os.environ.get("MMSEG_SWI_SLOTS", "3"))   # 2 -> 383 ms, 3 -> 378 ms per volume
        self._state = None
        self._slots = None
        self._slots_key = None
        self._weights_version = None
        self.launches_last = 0

    # -- helpers
    def _backbone(self):
        m = self.model
        return m.backbone if hasattr(m, "backbone") else m

    def device_volume(self, shape, device) -> Tensor:
        """A persistent device staging buffer for host volumes of `shape` (keeps its address: graph-safe)."""
        v = getattr(self, "_dev_vol", None)
        if v is None or tuple(v.shape) != tuple(shape) or v.device != device:
            v = self._dev_vol = torch.zeros(tuple(shape), dtype=torch.float32, device=device)
        return v

    def _get_slots(self, device):
        """Batch slots, kept for the life of the inferer (per device and numeric mode): each has its own engine buffers,
        window-origin slot and CUDA stream, and all share slot 0's packed weights.  The forward of batch k+1
        (tensor-bound convs + HBM-bound norm kernels) runs concurrently with the tail / blend of batch k (measured:
        2, 3 and 4 slots are within 0.5 %).  Blends all run on the caller's stream in window order (deterministic)."""
        bb = self._backbone()
        key = (str(device), bb.numeric_mode)
        if self._slots is not None and self._slots_key == key:
            return self._slots
        eng = bb.engine()
        slots = []
        for j in range(self.n_slots):
            e = eng if j == 0 else type(eng)(eng.module, eng.mode, weights_from=eng)
            slots.append({"eng": e, "stream": torch.cuda.Stream(device=device), "starts_dev": None, "logits": None,
                          "graphs": {}, "launches": {}, "feat": {}, "ev_fwd": torch.cuda.Event(), "ev_blend": None})
        self._slots, self._slots_key = slots, key
        return slots

    def _check_weights(self) -> None:
        """Drop every captured graph when a parameter / norm buffer changed since capture (a replay would
        read the previous — possibly freed — packed weights)."""
        from ...engine import module_version
        ver = module_version(self._backbone())
        if ver != self._weights_version:
            if self._slots is not None:
                for slot in self._slots:
                    slot["graphs"].clear()
                    slot["launches"].clear()
            self._weights_version = ver

    def _setup(self, C: int, vol_shape, device, n_local: Optional[int] = None):
        bb = self._backbone()
        slots = self._get_slots(device)
        starts_n = None
        nb_req = self.engine_batch
        key = (C, tuple(vol_shape), str(device), bb.numeric_mode, nb_req, n_local)
        if self._state is not None and self._state["key"] == key:
            return self._state
        K_out = bb.out_channels
        VZ, VY, VX = vol_shape
        starts = window_starts(vol_shape, self.roi, self.overlap)
        tabs, floor = importance_tables(self.roi, self.mode, self.sigma_scale)
        n_mine = len(starts) if n_local is None else n_local
        nb = pick_engine_batch(n_mine) if nb_req is None else min(nb_req, max(1, n_mine))
        prev = self._state
        acc = None
        if prev is not None and tuple(prev["acc"].shape) == (K_out + 1, VZ, VY, VX) and prev["acc"].device == device:
            acc = prev["acc"]
        st = {
            "key": key, "eng": slots[0]["eng"], "starts": starts, "nb": nb, "K": K_out, "slots": slots,
            "wz": tabs[0].to(device), "wy": tabs[1].to(device), "wx": tabs[2].to(device), "floor": floor,
            # all window origins live on the device; each batch is a stream-ordered D2D copy into the fixed
            # `starts_dev` slot the (graph-captured) kernels read, so no host sync sits between batches
            "starts_all": torch.tensor(starts, dtype=torch.int32).to(device),
            # weighted-logit accumulator and the count map share one allocation: acc[:K] = out, acc[K] = count,
            # so a rank's partial result travels as ONE tensor in the sharded exchange
            "acc": acc if acc is not None else torch.empty((K_out + 1, VZ, VY, VX), dtype=torch.float32, device=device),
        }
        for slot in slots:
            if slot["starts_dev"] is None or slot["starts_dev"].shape[0] < nb:
                slot["starts_dev"] = torch.zeros((nb, 3), dtype=torch.int32, device=device)
                slot["graphs"].clear()      # the graphs baked the old origin slot's address in
                slot["launches"].clear()
        st["out"], st["count"] = st["acc"][:K_out], st["acc"][K_out]
        # out_conv fused into the blend (no logits tensor) when the 1x1 head and the window geometry allow it
        eng = slots[0]["eng"]
        oc = getattr(eng.module, "out_conv", None)
        f0 = eng.module.features[0]
        st["fused_head"] = bool(
            os.environ.get("MMSEG_SWI_FUSED_HEAD", "1") == "1" and oc is not None and tuple(oc.kernel_size) == (1, 1, 1)
            and K_out <= 8 and f0 % 8 == 0 and f0 <= 128 and self.roi[2] % 4 == 0 and VX % 4 == 0
            and self.roi[2] // 4 <= 128 and all(s_[2] % 4 == 0 for s_ in starts))
        if not st["fused_head"]:
            for slot in slots:
                if slot["logits"] is None or slot["logits"].shape[0] < nb or slot["logits"].shape[1] != K_out:
                    slot["logits"] = torch.empty((nb, K_out, *self.roi), dtype=torch.float32, device=device)
        self._state = st
        return st

    def _forward_batch(self, st, slot, volume: Tensor, n: int) -> None:
        """gather -> forward for the n windows whose origins are in the slot's starts_dev[:n] (current stream)."""
        rz, ry, rx = self.roi
        slot["eng"].gather_windows(volume, slot["starts_dev"], n, self.roi)
        if st["fused_head"]:
            # features only; the buffer is persistent per batch size (graph replays of full batches keep reading it)
            slot["feat"][n] = slot["eng"].forward_blocked(n, rz, ry, rx, None, device=volume.device)
        else:
            slot["eng"].forward_blocked(n, rz, ry, rx, slot["logits"][:n])

    def _blend_batch(self, st, slot, n: int) -> None:
        if st["fused_head"]:
            oc = slot["eng"].module.out_conv
            feat = slot["feat"][n]
            for j in range(n):
                K.swi_logits_blend(feat, 0, oc.in_channels, j, oc.weight, oc.bias, slot["starts_dev"][j], st["wz"],
                                   st["wy"], st["wx"], st["floor"], st["out"], st["count"])
            return
        logits = slot["logits"]
        for j in range(n):
            K.swi_blend(logits[j:j + 1], slot["starts_dev"][j], 1, st["wz"], st["wy"], st["wx"], st["floor"], st["out"],
                        st["count"], (-1, 0, 0, 0, 0, 0))

    def _run_batch(self, st, volume: Tensor, n: int) -> None:
        """One batch, eagerly, on the current stream with slot 0 (used by bench.py's per-kernel profile pass)."""
        slot = st["slots"][0]
        self._forward_batch(st, slot, volume, n)
        self._blend_batch(st, slot, n)

    @torch.no_grad()
    def accumulate(self, volume: Tensor, lo: int = 0, hi: Optional[int] = None, ready=None) -> None:
        """Accumulate windows [lo, hi) of `volume` [C, VZ, VY, VX] (fp32, CUDA) into out / count (zeroed first).
        ready: optional [(z_end, event), ...] in ascending z — the volume is valid below z_end once `event` has
        completed (slab-wise host-to-device upload overlapping the first windows); a batch waits for the first event
        that covers its windows instead of for the whole volume."""
        _lib.require_device()
        assert volume.dim() == 4 and volume.is_cuda and volume.dtype == torch.float32 and volume.is_contiguous()
        n_total = len(window_starts(volume.shape[1:], self.roi, self.overlap)) if (hi is not None or lo) else None
        n_local = None if n_total is None else (n_total if hi is None else hi) - lo
        st = self._setup(volume.shape[0], volume.shape[1:], volume.device, n_local)
        self._check_weights()
        starts = st["starts"]
        hi = len(starts) if hi is None else hi
        main = torch.cuda.current_stream(volume.device)
        st["acc"].zero_()
        nb = st["nb"]
        for slot in st["slots"]:
            slot["stream"].wait_stream(main)      # the volume (and anything else queued by the caller) is ready
        i, k = lo, 0
        ri = 0
        vkey = (volume.data_ptr(), tuple(volume.shape))
        while i < hi:
            n = min(nb, hi - i)
            slot = st["slots"][k % len(st["slots"])]
            if ready:
                zneed = max(s_[0] for s_ in starts[i:i + n]) + self.roi[0]
                while ri < len(ready) - 1 and ready[ri][0] < zneed:
                    ri += 1
                slot["stream"].wait_event(ready[ri][1])
            with torch.cuda.stream(slot["stream"]):
                if slot["ev_blend"] is not None:   # the previous user of this slot's logits / origins has been blended
                    slot["stream"].wait_event(slot["ev_blend"])
                slot["starts_dev"][:n].copy_(st["starts_all"][i:i + n], non_blocking=True)
                if self.use_graph:
                    gkey = (n, vkey, st["fused_head"])
                    g = slot["graphs"].get(gkey)
                    if g is None:
                        self._forward_batch(st, slot, volume, n)  # warm-up: allocates workspaces, packs weights
                        slot["stream"].synchronize()
                        g = torch.cuda.CUDAGraph()
                        l0 = K.LAUNCHES[0]
                        # thread_local: only this thread's calls are checked against the capture — a CUDA call from
                        # another host thread (a DataLoader pin-memory thread, NVML samplers, ...) must not invalidate it
                        with capture_guard(), torch.cuda.graph(g, stream=slot["stream"], capture_error_mode="thread_local"):
                            self._forward_batch(st, slot, volume, n)
                        slot["launches"][gkey] = K.LAUNCHES[0] - l0
                        if len(slot["graphs"]) >= 8:      # bounded: volumes of many shapes do not pile up graphs
                            slot["graphs"].pop(next(iter(slot["graphs"])))
                        slot["graphs"][gkey] = g
                        g.replay()
                    else:
                        g.replay()
                        K.LAUNCHES[0] += slot["launches"][gkey]
                else:
                    self._forward_batch(st, slot, volume, n)
                slot["ev_fwd"].record(slot["stream"])
            main.wait_event(slot["ev_fwd"])
            self._blend_batch(st, slot, n)
            if slot["ev_blend"] is None:
                slot["ev_blend"] = torch.cuda.Event()
            slot["ev_blend"].record(main)
            i += n
            k += 1

    @torch.no_grad()
    def finalize(self, normalize: bool = True, labels: bool = True, z0: int = 0, z1: Optional[int] = None):
        """out /= count (in place) and / or argmax -> uint8 over the axis-0 slab [z0, z1), in place (the kernel takes
        the class-plane stride, so a slab needs no packed copy)."""
        st = self._state
        VZ, VY, VX = st["out"].shape[1:]
        z1 = VZ if z1 is None else z1
        lab = None
        if z1 > z0:
            lab = torch.empty((z1 - z0, VY, VX), dtype=torch.uint8, device=st["acc"].device) if labels else None
            K.swi_finalize(st["out"], st["count"], normalize, lab, z0, z1)
        return (st["out"] if normalize else None), lab

    @torch.no_grad()
    def __call__(self, inputs: Tensor, return_labels: bool = False):
        """inputs [1, C, H, W, D] -> logits [1, K, H, W, D] (input dtype fp32) or uint8 labels [H, W, D].
        The logits are a VIEW of the inferer's persistent accumulator (overwritten by the next call): internal
        callers only — the public `sliding_window_inference` returns a fresh tensor."""
        assert inputs.dim() == 5 and inputs.shape[0] == 1, "engine path takes one volume at a time (trainer.py:357)"
        vol = inputs[0].contiguous().float()
        self.accumulate(vol)
        out, lab = self.finalize(normalize=not return_labels, labels=return_labels)
        return lab if return_labels else out.unsqueeze(0)

    # ------------------------------------------------------------------ multi-GPU: window chunks + one exchange
    @torch.no_grad()
    def run_sharded(self, volume: Tensor, group=None, want: str = "all", ready=None):
        """One volume over all ranks of `group`: rank r evaluates its contiguous chunk of the ordered window list,
        then ONE exchange moves every partial (weighted logits + count, K+1 channels) that falls into another rank's
        axis-0 slab to that owner, which adds them in rank order (deterministic) and finalises its slab in place.
        `volume` must hold valid data at least over this rank's input range (see `input_range`).
        want: "all"   — uint8 labels [VZ, VY, VX] on every rank (all-gather of the slabs);
              "rank0" — the full label map on rank 0 only (gather; other ranks return None);
              "slab"  — (labels of this rank's own slab, (z0, z1)) and no label collective at all."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        starts = window_starts(volume.shape[1:], self.roi, self.overlap)
        lo, hi = shard_windows(len(starts), world, rank)
        self.accumulate(volume, lo, hi, ready=ready)
        st = self._state
        VZ = st["acc"].shape[1]
        if world == 1:
            lab = self.finalize(normalize=False, labels=True)[1]
            return (lab, (0, VZ)) if want == "slab" else lab
        plan = exchange_plan(starts, self.roi[0], VZ, world)
        exchange_partials(st["acc"], plan, rank, world, group, cache=st.setdefault("xchg", {}))
        z0, z1 = plan["slabs"][rank]
        _, lab = self.finalize(normalize=False, labels=True, z0=z0, z1=z1)
        if want == "slab":
            return lab, (z0, z1)
        return gather_label_slabs(lab, plan["slabs"], st["acc"].shape[1:], rank, world, group, volume.device,
                                  dst=0 if want == "rank0" else None)

    def input_range(self, vol_shape, world: int, rank: int) -> Tuple[int, int]:
        """Axis-0 range of the input volume that rank's window chunk reads."""
        starts = window_starts(vol_shape, self.roi, self.overlap)
        lo, hi = shard_windows(len(starts), world, rank)
        return touched_range(starts, lo, hi, self.roi[0])


def exchange_plan(starts, roi0: int, size0: int, world: int):
    """Who sends which axis-0 range to whom: sends[src][dst] = (z0, z1) = touched(src) ∩ slab(dst), src != dst."""
    slabs = owned_slabs(starts, roi0, size0, world)
    touched = [touched_range(starts, *shard_windows(len(starts), world, r), roi0) for r in range(world)]
    sends = [[None] * world for _ in range(world)]
    for src in range(world):
        for dst in range(world):
            if src == dst:
                continue
            z0, z1 = max(touched[src][0], slabs[dst][0]), min(touched[src][1], slabs[dst][1])
            if z1 > z0:
                sends[src][dst] = (z0, z1)
    return {"slabs": slabs, "touched": touched, "sends": sends}


def exchange_partials(acc: Tensor, plan, rank: int, world: int, group=None, cache: Optional[dict] = None) -> None:
    """The single data-path exchange of sharded inference: point-to-point sends of overlap regions to their owner
    (NCCL over NVLink on GPUs, gloo in the CPU tests), then a rank-ordered add on the owner.

    No staging copies: plane p of the range acc[p, z0:z1] is contiguous, so it is sent as is (one P2P op per plane,
    all batched into one group call); the receive buffers persist in `cache`; the add is one strided kernel per sender
    on CUDA (kernels.swi_add_partial) — in rank order, so the sums are deterministic."""
    import torch.distributed as dist
    sends = plan["sends"]
    P = acc.shape[0]
    peer = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
    ops, recv_bufs = [], {}
    for src in range(world):
        rng = sends[src][rank]
        if rng is not None:
            shape = (P, rng[1] - rng[0], *acc.shape[2:])
            buf = cache.get(("recv", src, shape)) if cache is not None else None
            if buf is None:
                buf = torch.empty(shape, dtype=acc.dtype, device=acc.device)
                if cache is not None:
                    cache[("recv", src, shape)] = buf
            recv_bufs[src] = (rng, buf)
            for p in range(P):
                ops.append(dist.P2POp(dist.irecv, buf[p], peer(src), group))
    for dst in range(world):
        rng = sends[rank][dst]
        if rng is not None:
            for p in range(P):
                ops.append(dist.P2POp(dist.isend, acc[p, rng[0]:rng[1]], peer(dst), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for src in sorted(recv_bufs):  # fixed (rank) order -> deterministic sums
        (z0, z1), buf = recv_bufs[src]
        if acc.is_cuda:
            K.swi_add_partial(acc, z0, z1, buf)
        else:   # CPU (gloo) tests of the exchange logic
            acc[:, z0:z1] += buf


def gather_label_slabs(lab: Optional[Tensor], slabs, vol_shape, rank: int, world: int, group, device,
                       dst: Optional[int] = None) -> Optional[Tensor]:
    """The per-rank uint8 label slabs -> the full label volume: on every rank (dst None: all-gather, slabs padded to
    the largest) or on rank `dst` only (point-to-point sends of exactly the slab bytes; other ranks return None)."""
    import torch.distributed as dist
    VZ, VY, VX = vol_shape
    z0, z1 = slabs[rank]
    if dst is not None:
        peer = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
        if rank != dst:
            if z1 > z0:
                for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, lab, peer(dst), group)]):
                    w.wait()
            return None
        full = torch.empty((VZ, VY, VX), dtype=torch.uint8, device=device)
        ops = [dist.P2POp(dist.irecv, full[a:b], peer(r), group) for r, (a, b) in enumerate(slabs) if r != dst and b > a]
        if z1 > z0:
            full[z0:z1].copy_(lab)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return full
    zmax = max(b - a for a, b in slabs)
    mine = torch.zeros((zmax, VY, VX), dtype=torch.uint8, device=device)
    if z1 > z0:
        mine[:z1 - z0] = lab
    flat = torch.empty((world * zmax, VY, VX), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(flat, mine, group=group)
    allb = flat.view(world, zmax, VY, VX)
    full = torch.empty((VZ, VY, VX), dtype=torch.uint8, device=device)
    for r, (a, b) in enumerate(slabs):
        if b > a:
            full[a:b] = allb[r, :b - a]
    return full


def get_inferer(model, roi_size, overlap: float, mode: str, sigma_scale: float = 0.125,
                engine_batch: Optional[int] = None) -> SlidingWindowInferer:
    """The model's cached inferer for these settings.  The cache lives ON the model object (a plain attribute, not a
    registered submodule), so it is collected with the model and never pins another model's buffers; at most four
    settings are kept per model."""
    cache = model.__dict__.setdefault("_mmseg_inferers", {})
    key = (tuple(int(r) for r in roi_size), float(overlap), mode, float(sigma_scale), engine_batch)
    inf = cache.get(key)
    if inf is None:
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))
        inf = cache[key] = SlidingWindowInferer(model, roi_size, overlap, mode, sigma_scale, engine_batch=engine_batch)
    return inf


@torch.no_grad()
def predict_volume(model, image_host: Tensor, roi_size=(96, 96, 96), overlap: float = 0.5, mode: str = "constant",
                   engine_batch: Optional[int] = None, group=None, out_host: Optional[Tensor] = None,
                   gather: str = "all") -> Optional[Tensor]:
    """Host-to-host inference of one volume — the arithmetic of Trainer.predict (reference trainer.py:357-367):
    image [C, H, W, D] fp32 in (pinned) host memory -> uint8 label map [H, W, D] in host memory.

    Every step is stream-ordered: H2D copy of the input range this rank needs, sliding-window accumulation, the
    sharded exchange when torch.distributed is initialised with more than one rank, finalize + argmax, D2H of labels.
    gather (multi-rank only): "all" — every rank returns the full label map (all-gather + a full D2H per rank);
    "rank0" — rank 0 returns the full map, the others None (one gather, one D2H: the cheap way to a complete result);
    "slab" — every rank returns only its own axis-0 slab, copied into out_host[z0:z1] (no label collective)."""
    import torch.distributed as dist
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("predict_volume needs the model on a CUDA device (no CPU fallback)")
    assert image_host.dim() == 4 and image_host.dtype == torch.float32 and not image_host.is_cuda
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    inf = get_inferer(model, roi_size, overlap, mode, engine_batch=engine_batch)
    vol = inf.device_volume(image_host.shape, dev)
    z0, z1 = inf.input_range(image_host.shape[1:], world, rank)
    # slab-wise upload on a copy stream: contiguous per-channel slabs (plain async H2D copies, no host staging), one
    # event per slab, so the first windows start after ~1/8 of the transfer instead of after all 629 MB
    main = torch.cuda.current_stream(dev)
    cs = getattr(inf, "_copy_stream", None)
    if cs is None:
        cs = inf._copy_stream = torch.cuda.Stream(device=dev)
    cs.wait_stream(main)
    ready = []
    slab = max(int(roi_size[0]), -(-(z1 - z0) // 8))
    with torch.cuda.stream(cs):
        for zs in range(z0, z1, slab):
            ze = min(zs + slab, z1)
            for c in range(image_host.shape[0]):
                vol[c, zs:ze].copy_(image_host[c, zs:ze], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            ready.append((ze, ev))
    VZ, VY, VX = image_host.shape[1:]
    if world > 1:
        res = inf.run_sharded(vol, group, want=gather, ready=ready)
    else:
        inf.accumulate(vol, ready=ready)
        res = inf.finalize(normalize=False, labels=True)[1]
        if gather == "slab":
            res = (res, (0, VZ))
    main.wait_stream(cs)
    if res is None:                       # gather == "rank0" on another rank
        torch.cuda.current_stream().synchronize()
        return None
    if out_host is None:
        out_host = torch.empty((VZ, VY, VX), dtype=torch.uint8, pin_memory=True)
    if gather == "slab":
        lab, (a, b) = res
        if b > a:
            out_host[a:b].copy_(lab, non_blocking=True)
    else:
        out_host.copy_(res, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_host


def _pad_to_roi(inputs: Tensor, roi: Sequence[int], cval: float):
    nsp = inputs.dim() - 2
    pad = []
    for k in range(nsp - 1, -1, -1):
        diff = max(roi[k] - inputs.shape[k + 2], 0)
        half = diff // 2
        pad.extend([half, diff - half])
    if not any(pad):
        return inputs, None
    return torch.nn.functional.pad(inputs, pad, mode="constant", value=cval), pad


def sliding_window_inference(inputs: Tensor, roi_size, sw_batch_size: int, predictor: Callable, overlap: float = 0.25,
                             mode: str = "constant", sigma_scale: float = 0.125, padding_mode: str = "constant",
                             cval: float = 0.0, **kwargs) -> Tensor:
    """Signature-compatible with monai.inferers.sliding_window_inference as called by the reference.

    `sw_batch_size` is accepted for compatibility; the engine picks its own batch (the result is independent of it).
    Returns a FRESH tensor like MONAI does (the inferer's accumulator is reused by the next call)."""
    if not inputs.is_cuda:
        raise RuntimeError("mmseg_b200 sliding_window_inference runs on CUDA tensors only (no CPU fallback)")
    if padding_mode != "constant":
        raise NotImplementedError("only padding_mode='constant' (the reference's default) is implemented")
    bb = predictor.backbone if hasattr(predictor, "backbone") else predictor
    if not hasattr(bb, "engine"):
        raise NotImplementedError("predictor must be a mmseg_b200 drop-in model (UNet3D / DualEncoder)")
    roi = [int(r) if r and r > 0 else int(s) for r, s in zip(roi_size, inputs.shape[2:])]
    orig = inputs.shape[2:]
    inputs, pad = _pad_to_roi(inputs.float(), roi, cval)
    inf = get_inferer(predictor, roi, overlap, mode, sigma_scale)
    outs = [inf(inputs[b:b + 1]).clone() for b in range(inputs.shape[0])]
    out = torch.cat(outs) if len(outs) > 1 else outs[0]
    if pad is not None:
        nsp = len(orig)
        crop = [slice(None), slice(None)]
        for k in range(nsp):
            before = pad[2 * (nsp - 1 - k)]
            crop.append(slice(before, before + orig[k]))
        out = out[tuple(crop)].contiguous()
    return out
