#!/usr/bin/env python
"""Headline benchmark: sliding-window inference voxels/s (CT+PET) — BASELINE.json configs[2].

Workload: synthetic 2-channel 512x512x300 CT+PET volume, UNet3D (early fusion, features 32..512, random-init seed 0),
roi 96^3, overlap 0.5, Gaussian blend -> 600 windows (241.9 algorithmic TFLOP) -> uint8 label map.
A "step" is one whole volume.  N>1: the ordered window list is cut into N contiguous chunks (axis-0 slabs), one
exchange of the overlapping partial sums, labels all-gathered — strong scaling of one volume.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]      the CPU oracle port of the reference's path

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-resident throughput, e2e = through
predict_volume() with pinned HOST buffers (H2D + D2H inside the timed region), roofline for the dominant kernel
(conv3d_tc_kernel, tensor bound) and for the norm/activation kernel (HBM bound), cpu_baseline, clocks, gpu_launches.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

VOL = (512, 512, 300)
ROI = (96, 96, 96)
OVERLAP = 0.5
MODE = "gaussian"
FEATURES = [32, 64, 128, 256, 512]
GF_PER_WINDOW = 403.2          # UNet3D(2->8) forward at 96^3, conv-type layers, 2*MACs
CROP = (192, 192, 240)         # exactly 36 windows (3 x 3 x 4): the bounded CPU sample (~14 s) / parity sample


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synthetic_volume(shape=VOL, seed=1234) -> torch.Tensor:
    """CT = smooth field * 350 - 100 + N(0, 30) HU -> window [-450, 250] -> [0, 1];
    PET = exp(N(0,1)) * smooth blobs / max.  Smooth fields = trilinear upsampling (x8) of coarse Gaussian noise."""
    g = torch.Generator().manual_seed(seed)
    coarse = [max(2, (s + 7) // 8 + 1) for s in shape]
    up = lambda t: torch.nn.functional.interpolate(t[None, None], size=shape, mode="trilinear", align_corners=True)[0, 0]
    ct = up(torch.randn(coarse, generator=g)) * 350.0 - 100.0
    ct += torch.randn(shape, generator=g) * 30.0
    ct = (ct.clamp_(-450.0, 250.0) + 450.0) / 700.0
    blobs = up(torch.randn(coarse, generator=g)).clamp_(min=0.0)
    pet = torch.exp(torch.randn(shape, generator=g)) * blobs
    pet /= pet.max()
    return torch.stack([ct, pet]).contiguous()


def model_config(device: str):
    return {"model": {"name": "unet", "in_channels": 2, "out_channels": 8,
                      "backbone": {"features": FEATURES, "norm": "instance"},
                      "fusion": {"type": "early"}, "head": {"dropout": 0.0}},
            "data": {"modalities": ["CT", "PET"]}, "hardware": {"device": device}}


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that clocks were not sampled
            self.nv, self.err = None, repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if self.nv is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["not_sampled"]}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU legs
REF_CROP = (192, 192, 144)     # exactly 18 windows (3 x 3 x 2): one step of the reference arm (>= 18)
HEADLINE_DEFAULT = "fp16i"     # fastest mode that met all four north_star gates on B200 (see `parity` in the bench line)


def plain_unet_state_dict(features=FEATURES, cin=2, cout=8):
    """Random-init UNet3D weights drawn with plain torch modules in the reference's construction order
    (unet.py:136-163: init_conv, encoders, decoders (up, conv), out_conv), i.e. the same RNG stream as
    torch.manual_seed(0); reference build_model(...) — used only when baseline/_ref is not available."""
    import torch.nn as nn
    sd = {}

    def block(name, ci, co):
        for j, (a, b) in enumerate(((ci, co), (co, co)), 1):
            c = nn.Conv3d(a, b, 3, padding=1)
            sd[f"backbone.{name}.conv{j}.weight"], sd[f"backbone.{name}.conv{j}.bias"] = c.weight.detach(), c.bias.detach()

    block("init_conv", cin, features[0])
    for i in range(len(features) - 1):
        block(f"encoders.{i}.conv", features[i], features[i + 1])
    for j, i in enumerate(range(len(features) - 1, 0, -1)):
        up = nn.ConvTranspose3d(features[i], features[i] // 2, 2, stride=2)
        sd[f"backbone.decoders.{j}.up.weight"], sd[f"backbone.decoders.{j}.up.bias"] = up.weight.detach(), up.bias.detach()
        block(f"decoders.{j}.conv", features[i], features[i - 1])
    oc = nn.Conv3d(features[0], cout, 1)
    sd["backbone.out_conv.weight"], sd["backbone.out_conv.bias"] = oc.weight.detach(), oc.bias.detach()
    return sd


def cpu_predictor(sd=None):
    """(predictor, kind, state_dict): the reference's own modules from baseline/_ref when installed (kind "reference":
    stock build_model + forward, seed-0 weights unless `sd` is given), else the oracle port of the forward (kind "port")."""
    from oracle import install_ref
    if install_ref.available():
        build_model, _ = install_ref.import_reference()
        torch.manual_seed(0)
        m = build_model(model_config("cpu")).eval()
        if sd is not None:
            m.load_state_dict(sd, strict=True)
        return (lambda w: m(w)), "reference", {k: v.detach().clone() for k, v in m.state_dict().items()}
    from oracle.models import unet3d_forward
    if sd is None:
        torch.manual_seed(0)
        sd = plain_unet_state_dict()
    return (lambda w: unet3d_forward(sd, w)), "port", sd


def cpu_sample(predictor, vol_cpu, crop, threads):
    """The reference's inference loop on the CPU over a crop: MONAI's sliding_window_inference as restated in
    oracle/sliding_window.py (MONAI itself is installed nowhere) around `predictor`, sw_batch_size 4 as in
    trainer.py:386-392.  Returns (seconds, n_windows, logits)."""
    from oracle.sliding_window import sliding_window_inference as oswi, window_starts
    torch.set_num_threads(threads)
    x = vol_cpu[:, :crop[0], :crop[1], :crop[2]].unsqueeze(0).contiguous()
    n_win = len(window_starts(crop, ROI, OVERLAP))
    t0 = time.perf_counter()
    with torch.no_grad():
        out = oswi(x, ROI, 4, predictor, overlap=OVERLAP, mode=MODE)
    return time.perf_counter() - t0, n_win, out


def full_volume_equiv(seconds, n_win):
    """voxels/s of the full 512x512x300 job extrapolated from `n_win` windows (600 windows per volume)."""
    from oracle.sliding_window import window_starts
    total = len(window_starts(VOL, ROI, OVERLAP))
    return (VOL[0] * VOL[1] * VOL[2]) / (seconds / n_win * total)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads:
    stock `build_model` + `forward` from baseline/_ref (installed by oracle/install_ref.py; the oracle port of the
    forward if that copy is absent) inside the restated MONAI sliding-window loop.  None of this repo's models, kernels
    or engine is imported.  One step = 18 of the 600 windows (192x192x144 crop), extrapolated.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    predictor, kind, _ = cpu_predictor()
    vol = synthetic_volume(REF_CROP, 1234)
    times = []
    n_win = 0
    for i in range(args.warmup + args.steps):
        t, n_win, _ = cpu_sample(predictor, vol, REF_CROP, threads)
        if i >= args.warmup:
            times.append(t)
        log(f"[reference:{kind}] step {i}: {t:.2f} s for {n_win} windows")
    ms = sum(times) / len(times) * 1e3
    v = full_volume_equiv(ms / 1e3, n_win)
    sample = (f"{n_win} of the 600 windows per step ({REF_CROP[0]}x{REF_CROP[1]}x{REF_CROP[2]} crop, sw_batch 4), extrapolated "
              f"x600/{n_win} to the full volume; model forward = " +
              ("the reference's own modules (baseline/_ref, stock build_model)" if kind == "reference" else "oracle port") +
              "; sliding-window loop = oracle/sliding_window.py (MONAI restated; MONAI is not installed)")
    cfg = workload_config(args, 1)
    cfg["numeric_mode"] = "fp32 (PyTorch CPU)"
    emit({
        "impl": "reference", "metric": "sliding-window inference voxels/s (CT+PET)", "value": v, "unit": "voxels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, world):
    return {"workload": "sliding_window_inference 2ch CT+PET 512x512x300, UNet3D early fusion features 32-512, "
                        "roi 96^3 overlap 0.5 gaussian blend, 600 windows -> uint8 labels (BASELINE.json configs[2])",
            "windows": 600, "engine_batch": args.engine_batch or "auto", "numeric_mode": args.mode,
            "parallelism": f"window-chunks x{world} (axis-0 slabs, one partial-sum exchange)" if world > 1 else "single GPU",
            "l2": "inputs_exceed_l2 (volume 629 MB + accumulator 2.8 GB per step >> 126 MB L2)"}


GATE_LIMITS = {"max_abs": 2e-2, "rel_l2": 1e-3, "label_agreement": 0.999, "dice_abs_err": 1e-3}


def parity_metrics(got, ref_logits):
    """The four north_star gates of one mode against the fp32 CPU reference on the crop."""
    ref_lab = ref_logits.argmax(1)[0]
    lab = got.argmax(1)[0]
    dices = []
    for c in range(1, ref_logits.shape[1]):
        a, b = (lab == c), (ref_lab == c)
        dices.append((2.0 * (a & b).sum().item() + 1e-5) / (a.sum().item() + b.sum().item() + 1e-5))
    d = (got - ref_logits).double()
    m = {"max_abs": d.abs().max().item(), "rel_l2": (d.norm() / ref_logits.double().norm()).item(),
         "label_agreement": (lab == ref_lab).double().mean().item(), "dice_vs_ref_mean_fg": sum(dices) / len(dices)}
    m["passes_gates"] = bool(m["max_abs"] <= GATE_LIMITS["max_abs"] and m["rel_l2"] <= GATE_LIMITS["rel_l2"]
                             and m["label_agreement"] >= GATE_LIMITS["label_agreement"]
                             and abs(1.0 - m["dice_vs_ref_mean_fg"]) <= GATE_LIMITS["dice_abs_err"])
    return m


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a B200; there is no CPU fallback (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import mmseg_b200  # noqa: F401
    from mmseg_b200 import kernels as K
    from mmseg_b200.numerics import LADDER, MODES
    from mmseg_b200.src.models.build import build_model
    from mmseg_b200.src.trainer.inference import SlidingWindowInferer, predict_volume, shard_windows, get_inferer

    torch.manual_seed(0)
    model = build_model(model_config("cuda")).eval()
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    headline = args.mode if args.mode != "auto" else HEADLINE_DEFAULT

    t0 = time.time()
    vol_host = synthetic_volume().pin_memory()
    log(f"[rank {rank}] synthetic volume {tuple(vol_host.shape)} in {time.time() - t0:.1f} s")
    nvox = VOL[0] * VOL[1] * VOL[2]
    n_windows = 600
    eb = args.engine_batch or None          # None: per-rank choice (pick_engine_batch: full, graph-replayed batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = K.LAUNCHES[0]
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        launches = torch.tensor([K.LAUNCHES[0] - l0], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        return ms.item() / steps, int(launches.item()), out

    # ---- CPU reference on the bounded sample + the parity ladder (rank 0 at N=1 only)
    cpu_baseline, ladder, ref_logits = None, {}, None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        predictor, kind, _ = cpu_predictor(sd_cpu)
        t, n_win, ref_logits = cpu_sample(predictor, vol_host, CROP, threads)
        cpu_baseline = {"value": full_volume_equiv(t, n_win), "unit": "voxels/s", "cores": threads, "kind": kind,
                        "sample": f"{n_win} of the 600 windows ({CROP[0]}x{CROP[1]}x{CROP[2]} crop of the same volume, "
                                  f"{t:.1f} s), extrapolated x600/{n_win}; forward = " +
                                  ("the reference's own modules (baseline/_ref)" if kind == "reference" else "oracle port") +
                                  ", sliding-window loop = MONAI restated (oracle/sliding_window.py)"}
        crop_dev = vol_host[:, :CROP[0], :CROP[1], :CROP[2]].contiguous().to(dev)
        for mode in LADDER:
            model.set_numeric_mode(mode)
            ci = SlidingWindowInferer(model, ROI, OVERLAP, MODE, engine_batch=6, use_graph=False)
            got = ci(crop_dev.unsqueeze(0)).cpu()
            ladder[mode] = {"dtype": MODES[mode].bench_dtype, "mma_passes": MODES[mode].passes, **parity_metrics(got, ref_logits)}
            del ci, got
            log(f"[ladder] {mode}: " + json.dumps(ladder[mode]))
        del crop_dev
        torch.cuda.empty_cache()

    # ---- throughput of every rung at N=1 (short), then the headline mode with the requested steps
    def measure_mode(mode, steps, warmup, with_e2e):
        model.set_numeric_mode(mode)
        inf = get_inferer(model, ROI, OVERLAP, MODE, engine_batch=eb)   # predict_volume() reuses this engine
        vol_dev = inf.device_volume(vol_host.shape, dev)
        vol_dev.copy_(vol_host)

        def step_resident():
            if world > 1:
                return inf.run_sharded(vol_dev, want="all")
            inf.accumulate(vol_dev)
            return inf.finalize(normalize=False, labels=True)[1]

        ms_step, launches, labels = timed(step_resident, steps, warmup)
        res = {"ms_per_step": ms_step, "value": nvox / (ms_step * 1e-3), "launches": launches}
        if with_e2e:
            # end to end through the public host API (pinned host in, pinned host out): rank 0 receives the full map
            out_host = torch.empty(VOL, dtype=torch.uint8).pin_memory()
            e2e_steps = max(1, min(steps, 5))
            ms_e2e, _, _ = timed(lambda: predict_volume(model, vol_host, ROI, OVERLAP, MODE, eb, out_host=out_host,
                                                        gather="rank0"), e2e_steps, 1)
            if rank == 0:
                assert torch.equal(out_host, labels.cpu()), "end-to-end labels differ from the device-resident run"
            res["e2e_ms"], res["e2e_steps"] = ms_e2e, e2e_steps
        res["inf"], res["labels"] = inf, labels
        return res

    if world == 1 and not args.no_ladder:
        for mode in LADDER:
            r = measure_mode(mode, 2, 1, with_e2e=True)
            ladder.setdefault(mode, {"dtype": MODES[mode].bench_dtype, "mma_passes": MODES[mode].passes})
            ladder[mode].update({"ms_per_step": r["ms_per_step"], "value": r["value"],
                                 "e2e_value": nvox / (r["e2e_ms"] * 1e-3), "steps": 2, "warmup": 1})
            log(f"[ladder] {mode}: {r['ms_per_step']:.1f} ms/volume -> {r['value'] / 1e6:.1f} Mvox/s")
            r["inf"]._state = None
            del r
            model.__dict__.pop("_mmseg_inferers", None)
            torch.cuda.empty_cache()
        if args.mode == "auto" and ladder and all("passes_gates" in v for v in ladder.values()):
            passing = [m for m in LADDER if ladder[m]["passes_gates"]]
            if passing:   # the headline rule: the fastest (measured) mode among those that meet every gate
                headline = min(passing, key=lambda m: ladder[m]["ms_per_step"])

    sampler = ClockSampler(local)
    sampler.start()
    hr = measure_mode(headline, args.steps, args.warmup, with_e2e=True)
    sampler.stop_flag = True
    sampler.join()
    inf, labels = hr["inf"], hr["labels"]
    ms_step, launches, value = hr["ms_per_step"], hr["launches"], hr["value"]
    ms_e2e = hr["e2e_ms"]
    log(f"[rank {rank}] headline mode {headline}: resident {ms_step:.2f} ms/volume -> {value / 1e6:.1f} Mvox/s; "
        f"e2e {ms_e2e:.2f} ms -> {nvox / (ms_e2e * 1e-3) / 1e6:.1f} Mvox/s")
    z0, z1 = inf.input_range(VOL, world, rank)
    h2d = torch.tensor([2 * (z1 - z0) * VOL[1] * VOL[2] * 4], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(h2d)
    e2e = {"value": nvox / (ms_e2e * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": int(h2d.item()),
           "d2h_bytes_per_step": nvox,   # the gathered uint8 label map is read back once, by rank 0
           "ms_per_step": ms_e2e, "steps": hr["e2e_steps"],
           "api": "predict_volume(model, pinned_host_volume, gather='rank0') -> pinned uint8 labels (what Trainer.predict_array calls)"}
    if ladder.get(headline) is not None:
        ladder[headline].update({"ms_per_step": ms_step, "value": value, "e2e_value": e2e["value"],
                                 "steps": args.steps, "warmup": args.warmup})

    # ---- N > 1: does the sharded result match the single-GPU one?  Rank 0 recomputes the whole volume alone.
    agreement_vs_n1 = None
    if world > 1 and not args.no_selfcheck:
        if rank == 0:
            # same engine batch as the sharded run: a window's logits depend on the batch size only through the tile plan
            # (summation order of the InstanceNorm statistics), so what is left is the order of the overlap adds
            solo = SlidingWindowInferer(model, ROI, OVERLAP, MODE, engine_batch=inf._state["nb"])
            lab1 = solo(inf._dev_vol.unsqueeze(0), return_labels=True)
            agreement_vs_n1 = (lab1 == labels).double().mean().item()
            log(f"[selfcheck] sharded x{world} labels vs single-GPU labels: {agreement_vs_n1 * 100:.5f}% identical")
            del solo, lab1
        dist.barrier()

    def secondary_train(which):
        """training samples/s (BASELINE.json configs[1] / configs[4]); never allowed to break the headline line"""
        if args.no_train:
            return None
        try:
            inf._state = None
            inf._dev_vol = None
            model.__dict__.pop("_mmseg_inferers", None)
            torch.cuda.empty_cache()
            return train_measure(3, 2, world, rank, dev, use_graph=True, cpu=(world == 1 and not args.no_cpu), which=which)
        except Exception as e:
            torch.cuda.synchronize()
            return {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        secondary_train("train")
        secondary_train("train5")
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: one batch of windows, every C-ABI launch bracketed by CUDA events on the launching stream
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_src = "fallback (B200_PROFILING.md: 1590 TFLOP/s burst, 1400 sustained, 6650 GB/s)"
    tf_peak, hbm_peak = 1400.0, 6650.0
    if os.path.exists(pk):
        peaks = json.load(open(pk))
        tf_peak, hbm_peak = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json: bf16_tflops_sustained, hbm_gbs)"

    def profile_batch(mode):
        """Per-kernel CUDA-event profile of one 8-window batch in `mode` (eager, serialised on one stream)."""
        model.set_numeric_mode(mode)
        pi = SlidingWindowInferer(model, ROI, OVERLAP, MODE, engine_batch=8, use_graph=False)
        vd = inf._dev_vol
        st = pi._setup(vd.shape[0], vd.shape[1:], vd.device)
        pi._check_weights()
        nb = st["nb"]
        st["acc"].zero_()
        st["slots"][0]["starts_dev"][:nb].copy_(st["starts_all"][:nb])
        pi._run_batch(st, vd, nb)
        torch.cuda.synchronize()
        K.PROFILE = []
        for _ in range(3):
            pi._run_batch(st, vd, nb)
        torch.cuda.synchronize()
        prof, K.PROFILE = K.PROFILE, None
        agg, layers = {}, {}
        for name, info, a, b in prof:
            ms = a.elapsed_time(b)
            d = agg.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "issued": 0.0, "bytes": 0.0})
            d["ms"] += ms
            d["n"] += 1
            if info:
                d["flops"] += info.get("flops", 0.0)
                d["issued"] += info.get("issued_flops", 0.0)
                d["bytes"] += info.get("bytes", 0.0)
                L = layers.setdefault((name, info["layer"]), {"ms": 0.0, "n": 0, "info": info})
                L["ms"] += ms
                L["n"] += 1
        total_ms = sum(d["ms"] for d in agg.values())
        conv, norm = agg["mmseg_conv3d_fwd"], agg["mmseg_instnorm_act_apply"]
        conv_tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        issued_tf = conv["issued"] / (conv["ms"] * 1e-3) / 1e12
        norm_gbs = norm["bytes"] / (norm["ms"] * 1e-3) / 1e9
        log(f"[{mode}] per-kernel shares of one {nb}-window batch (CUDA events, 3 repeats):")
        for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            log(f"  {name:28s} {d['ms'] / 3:8.3f} ms  {100 * d['ms'] / total_ms:5.1f}%  launches {d['n'] // 3}")
        log("conv layers:")
        for (name, layer), L in sorted(layers.items(), key=lambda kv: -kv[1]["ms"]):
            i = L["info"]
            if name == "mmseg_conv3d_fwd":
                log(f"  {layer:36s} {L['ms'] / L['n']:7.3f} ms  {i['flops'] / (L['ms'] / L['n'] * 1e-3) / 1e12:7.1f} TF/s  "
                    f"tile {i['tile']} ctas {i['ctas']}")
            else:
                log(f"  norm {layer:31s} {L['ms'] / L['n']:7.3f} ms  {i['bytes'] / (L['ms'] / L['n'] * 1e-3) / 1e9:7.0f} GB/s")
        passes = MODES[mode].passes
        roof = {"bound": "tensor", "kernel": "conv3d_tc_kernel + conv3d_roll_kernel", "numeric_mode": mode,
                "achieved": conv_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": conv_tf / tf_peak,
                "mma_passes": passes, "issued_tflops": issued_tf, "frac_issued": issued_tf / tf_peak,
                "frac_of_nominal_2250": conv_tf / 2250.0, "peak_source": peak_src,
                "launches": conv["n"] // 3, "avg_launch_ms": conv["ms"] / conv["n"], "share_of_step": conv["ms"] / total_ms,
                "note": "achieved = ALGORITHMIC conv FLOPs (2*vox*Cin*Cout*k^3, counted once whatever the number of MMA "
                        "passes) of one %d-window batch / sum of conv launch durations; issued_tflops = tensor-pipe work "
                        "actually issued (x%d passes, incl. K / N padding)" % (nb, passes)}
        roof_norm = {"bound": "hbm", "kernel": "instnorm_apply(_pool)_kernel", "numeric_mode": mode, "achieved": norm_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": norm_gbs / hbm_peak, "frac_of_nominal_8000": norm_gbs / 8000.0,
                     "traffic": None, "share_of_step": norm["ms"] / total_ms, "avg_launch_ms": norm["ms"] / norm["n"]}
        del pi
        return roof, roof_norm

    roofline, roofline_norm = profile_batch(headline)
    tj = os.path.join(ROOT, "profiles", "r02_conv_traffic.json")
    roofline["traffic"], roofline["traffic_note"] = None, "no ncu --set full capture committed for this mode"
    if os.path.exists(tj):
        tdat = json.load(open(tj))
        if tdat.get("numeric_mode") == headline:
            roofline["traffic"] = tdat.get("dram_bytes_per_launch_avg")
            roofline["traffic_note"] = tdat.get("note")
    roofline_fast = None
    if world == 1 and headline != "bf16" and not args.no_ladder:
        roofline_fast, _ = profile_batch("bf16")
    model.set_numeric_mode(headline)
    step_tf = n_windows * GF_PER_WINDOW / 1e3 / (ms_step * 1e-3) / world

    train_res = secondary_train("train")
    train5_res = secondary_train("train5")
    attn_res = None
    if world == 1 and not args.no_train:
        try:
            attn_res = attention_measure(dev)
        except Exception as e:   # never allowed to break the headline line
            torch.cuda.synchronize()
            attn_res = {"error": f"{type(e).__name__}: {e}"}

    swin_res = None
    if world == 1 and not args.no_train:
        try:
            swin_res = swin_measure(dev, cpu=not args.no_cpu)
        except Exception as e:   # never allowed to break the headline line
            torch.cuda.synchronize()
            swin_res = {"error": f"{type(e).__name__}: {e}"}

    cfg = workload_config(args, world)
    cfg["numeric_mode"] = headline
    cfg["engine_batch"] = inf._state["nb"] if inf._state is not None else (args.engine_batch or "auto")
    line = {
        "metric": "sliding-window inference voxels/s (CT+PET)", "value": value, "unit": "voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": MODES[headline].bench_dtype,
        "data": "synthetic", "config": cfg,
        "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
        "roofline": roofline, "roofline_norm": roofline_norm, "roofline_fast_mode": roofline_fast,
        "step_tflops_per_gpu": step_tf, "step_frac_of_tensor_peak": step_tf / tf_peak,
        "cpu_baseline": cpu_baseline,
        "headline_mode": headline,
        "headline_rule": "fastest numeric mode of the ladder that meets all four north_star gates on the 36-window crop "
                         "(max-abs <= 2e-2, rel-L2 <= 1e-3, labels >= 99.9 %, Dice within 1e-3 of the reference)",
        "parity": ladder or None, "gates": GATE_LIMITS,
        "label_agreement_vs_n1": agreement_vs_n1,
        "train": train_res, "train5": train5_res, "attention": attn_res, "swin": swin_res,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ training workload
TRAIN_S = 128
# SURVEY.md §8(d) algorithmic FLOPs per train step and GPU: 3 x forward minus the first-layer dgrads
TRAIN_WORKLOADS = {
    # BASELINE.json configs[1]: DualEncoder cross-attention (== mean in the reference) CT+PET 128^3, batch 2/GPU
    "train": dict(M=2, B=2, fusion="cross_attention", gf_per_step=7394.0,
                  label="DualEncoder(fusion=cross_attention -> mean) 2 modalities 128^3, batch 2/GPU (BASELINE.json configs[1])"),
    # BASELINE.json configs[4]: 4 modalities (CT/PET/MRI/US), attention-gate fusion, 128^3, batch 4/GPU
    "train5": dict(M=4, B=4, fusion="attention", gf_per_step=3.0 * 7200.5 - 4 * 14.5,   # four first-layer dgrads (Cin=1, B=4) never run
                   label="DualEncoder(fusion=attention gate) 4 modalities CT/PET/MRI/US 128^3, batch 4/GPU (BASELINE.json configs[4])"),
}


def train_config(device, M=2, fusion="cross_attention"):
    return {"model": {"name": "dual_encoder", "in_channels": M, "out_channels": 8,
                      "backbone": {"features": FEATURES, "norm": "instance"},
                      "fusion": {"type": fusion}, "head": {"dropout": 0.1}},
            "data": {"modalities": ["CT", "PET", "MRI", "US"][:M]},
            "training": {"epochs": 1, "optimizer": {"name": "adamw", "lr": 1e-4, "weight_decay": 1e-5},
                         "loss": {"name": "dice_ce", "dice_weight": 0.5, "ce_weight": 0.5}, "accumulation_steps": 1},
            "hardware": {"device": device, "mixed_precision": True, "cuda_graph": True},
            "experiment": {"output_dir": "/tmp/mmseg_b200_bench", "name": "train"}}


def train_measure(steps, warmup, world, rank, dev, use_graph=True, cpu=False, which="train"):
    """BASELINE.json configs[1] ("train") / configs[4] ("train5"): DualEncoder on 128^3 patches, bf16 kernels, DiceCE,
    AdamW; data-parallel over `world` ranks (weak scaling).  Returns a dict with device-resident and end-to-end (pinned
    host batch in, loss out) samples/s, a SERIALISED per-kernel profile, and (N=1, cpu=True) the full-size loss parity
    against the CPU oracle's batch-1 step."""
    import torch.distributed as dist
    import mmseg_b200  # noqa: F401
    from mmseg_b200 import kernels as K
    from mmseg_b200.src.models.build import build_model
    from mmseg_b200.src.trainer.trainer import Trainer
    torch.cuda.reset_peak_memory_stats(dev)   # peak_mem_gb since here (live buffers of an earlier workload included)
    W = TRAIN_WORKLOADS[which]
    M, TRAIN_B = W["M"], W["B"]
    torch.manual_seed(0)                       # identical replicas
    cfg = train_config("cuda", M, W["fusion"])
    model = build_model(cfg)
    tr = Trainer(cfg, model)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand((TRAIN_B, M, TRAIN_S, TRAIN_S, TRAIN_S), generator=g).pin_memory()
    y_host = torch.randint(0, 8, (TRAIN_B, TRAIN_S, TRAIN_S, TRAIN_S), generator=g).pin_memory()
    x, y = x_host.to(dev), y_host.to(dev)
    tr.model.train()
    mode = "eager"
    step = lambda a, b: tr.train_step(a, b)
    if use_graph:
        try:
            step = tr.graphed_train_step(x, y)
            mode = "cuda_graph"
        except Exception as e:  # capture can fail (e.g. a collective that is not capturable): measure eagerly
            log(f"[rank {rank}] CUDA-graph capture of the train step failed ({type(e).__name__}: {e}); running eagerly")
            torch.cuda.synchronize()
            step = lambda a, b: tr.train_step(a, b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / n, out

    ms, loss = timed(lambda: step(x, y), steps, warmup)

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        return step(xd, yd).item()          # device->host read of the loss, like trainer.py:260
    ms_e2e, last = timed(e2e_step, max(1, min(steps, 5)), 1)
    res = {"metric": f"training samples/s (DualEncoder {M} modalities 128^3, batch {TRAIN_B}/GPU, bf16 step)",
           "value": world * TRAIN_B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "n_gpus": world,
           "scaling": "weak", "mode": mode, "loss": float(loss.item() if torch.is_tensor(loss) else loss),
           "e2e": {"value": world * TRAIN_B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": world * (x_host.numel() * 4 + y_host.numel() * 8), "d2h_bytes_per_step": world * 4},
           "tflops_per_gpu": W["gf_per_step"] / 1e3 / (ms * 1e-3), "gpu_launches_per_step": None,
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
           "config": {"workload": W["label"] + ", features 32-512, dropout 0.1, DiceCE, AdamW, fwd+bwd+allreduce+step",
                      "parallelism": f"dp{world}"}}
    # one eager profiled step with the weight-gradient side stream switched OFF, so every kernel is bracketed alone on
    # one stream (kernel_ms is serialised kernel time, not time under overlap); also counts the launches of a step
    os.environ["MMSEG_WGRAD_SIDE_STREAM"] = "0"
    try:
        tr.train_step(x, y)
        torch.cuda.synchronize()
        K.PROFILE = []
        l0 = K.LAUNCHES[0]
        tr.train_step(x, y)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MMSEG_WGRAD_SIDE_STREAM", None)
    res["gpu_launches_per_step"] = K.LAUNCHES[0] - l0
    prof, K.PROFILE = K.PROFILE, None
    agg = {}
    for name, info, a, b in prof:
        d = agg.setdefault(name, [0.0, 0.0, 0.0])
        d[0] += a.elapsed_time(b)
        if info:
            d[1] += info.get("flops", 0.0)
            d[2] += info.get("bytes", 0.0)
    tot = sum(d[0] for d in agg.values())
    res["kernel_ms"] = {k: round(v[0], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    res["kernel_ms_note"] = "serialised: one eager step, weight gradients on the main stream, each C-ABI call bracketed by CUDA events"
    res["kernel_ms_total"] = round(tot, 3)
    res["kernel_tflops"] = {k: round(v[1] / v[0] / 1e9, 1) for k, v in agg.items() if v[1] > 0}
    res["kernel_gbs"] = {k: round(v[2] / v[0] / 1e6, 0) for k, v in agg.items() if v[2] > 0}
    if rank == 0:
        log(f"{which} step {mode}: {ms:.2f} ms -> {res['value']:.1f} samples/s ({res['tflops_per_gpu']:.0f} TFLOP/s/GPU); "
            f"e2e {ms_e2e:.2f} ms; kernels (serialised) {tot:.2f} ms: " +
            ", ".join(f"{k.replace('mmseg_', '')} {v:.2f}" for k, v in list(res['kernel_ms'].items())[:8]))
    if cpu and rank == 0 and world == 1:
        # full-size parity of the training forward: the kernel path's loss on sample 0 (dropout off) against the CPU
        # oracle's batch-1 step on the same weights (fp32 autograd restatement of trainer.py:250-253); gate 1e-3 relative
        from oracle.train import train_step as oracle_step
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k[len("backbone."):]: v.detach().cpu() for k, v in tr.model.state_dict().items()}
        bb = tr.model.backbone
        p_old = bb.dropout.p if hasattr(bb.dropout, "p") else None
        if p_old is not None:
            bb.dropout.p = 0.0
        try:
            with torch.enable_grad():
                loss_gpu = float(tr.criterion(tr.model(x[:1]), y[:1]).item())
        finally:
            if p_old is not None:
                bb.dropout.p = p_old
        t0 = time.perf_counter()
        loss_cpu, _, _ = oracle_step("dual", sd, dict(L=len(FEATURES), M=M, fusion=W["fusion"]), x_host[:1], y_host[:1], torch.float32)
        t = time.perf_counter() - t0
        res["loss_gpu_sample0"], res["loss_oracle_sample0"] = loss_gpu, float(loss_cpu)
        res["loss_rel_vs_oracle"] = abs(loss_gpu - loss_cpu) / abs(loss_cpu)
        res["cpu_baseline"] = {"value": 1.0 / t, "unit": "samples/s", "cores": threads, "kind": "port",
                               "sample": f"one fwd+DiceCE+bwd step at batch 1 ({t:.1f} s) through the oracle (CPU fp32 autograd)"}
        log(f"{which}: loss sample 0 kernels {loss_gpu:.6f} vs oracle {loss_cpu:.6f} (rel {res['loss_rel_vs_oracle']:.2e}); CPU step {t:.1f} s")
    del tr, model
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ attention leg
# (C, spatial side, CPU seconds of the reference module at B = 2 from BASELINE.md §2, 8 vCPU)
ATTN_SHAPES = [(512, 8, 1.30), (256, 16, 0.37), (128, 24, 4.56), (128, 32, None)]


def attention_measure(dev, steps=10, warmup=3):
    """CrossAttentionFusion forward (reference src/models/fusion/attention_fusion.py:120-164) at B = 2, 4 heads: the fused
    tcgen05 attention core (softmax(Q K^T / sqrt(hd)) V without the N x N matrix) and the whole module (q / kv / out
    projections on the conv kernel, residual + InstanceNorm).  TFLOP/s against the algorithmic 4*B*N^2*C of the core."""
    import mmseg_b200  # noqa: F401
    from mmseg_b200 import kernels as K
    from mmseg_b200.kernels import Blocked
    from mmseg_b200.src.models.fusion.attention_fusion import CrossAttentionFusion
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf_peak = json.load(open(pk))["bf16_tflops_sustained"] if os.path.exists(pk) else 1400.0
    out = []
    B = 2
    for C, S, cpu_s in ATTN_SHAPES:
        torch.manual_seed(0)
        m = CrossAttentionFusion(C, num_heads=4).to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(1)
        q_in, kv_in, dst = (Blocked(B, C, S, S, S, False, dev) for _ in range(3))
        K.pack_ncdhw(torch.randn((B, C, S, S, S), device=dev, generator=g), q_in)
        K.pack_ncdhw(torch.randn((B, C, S, S, S), device=dev, generator=g), kv_in)
        with torch.no_grad():
            for _ in range(warmup):
                m.forward_blocked(q_in, kv_in, dst)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                m.forward_blocked(q_in, kv_in, dst)
            e1.record()
            torch.cuda.synchronize()
            ms_module = e0.elapsed_time(e1) / steps
            K.PROFILE = []
            for _ in range(steps):
                m.forward_blocked(q_in, kv_in, dst)
            torch.cuda.synchronize()
            prof, K.PROFILE = K.PROFILE, None
        core = [a.elapsed_time(b) for name, _, a, b in prof if name == "mmseg_cross_attention_fwd"]
        ms_core = sum(core) / len(core)
        N = S ** 3
        flops = 4.0 * B * N * N * C
        proj = 8.0 * B * N * C * C
        row = {"C": C, "tokens": N, "batch": B, "heads": 4, "head_dim": C // 4, "core_ms": ms_core,
               "core_tflops": flops / (ms_core * 1e-3) / 1e12, "core_frac_of_tensor_peak": flops / (ms_core * 1e-3) / 1e12 / tf_peak,
               "module_ms": ms_module, "module_tflops": (flops + proj) / (ms_module * 1e-3) / 1e12,
               "reference_cpu_s_8vcpu": cpu_s, "speedup_vs_reference_cpu": (cpu_s / (ms_module * 1e-3)) if cpu_s else None}
        out.append(row)
        log(f"[attention] C={C} N={N}: core {ms_core:.3f} ms = {row['core_tflops']:.0f} TFLOP/s "
            f"({100 * row['core_frac_of_tensor_peak']:.0f}% of sustained peak), module {ms_module:.3f} ms"
            + (f", reference CPU {cpu_s:.2f} s" if cpu_s else ""))
        del m, q_in, kv_in, dst
        torch.cuda.empty_cache()
    return {"metric": "CrossAttentionFusion forward (B=2, 4 heads), algorithmic 4*B*N^2*C", "peak_tflops": tf_peak, "shapes": out}


def swin_measure(dev, steps=5, warmup=3, cpu=True):
    """SwinUNETR feature_size 48, 2-channel 96^3 patches (BASELINE.json configs[3]), FORWARD through the drop-in module
    (the backward of this model is not built): ms per batch, patches/s, algorithmic TFLOP/s (convs + linear layers +
    window attention), the per-kernel split of one eager forward, and parity against the CPU oracle restatement
    (PARITY UNPINNED: MONAI is absent, see oracle/swin_unetr.py) whose run time is the CPU baseline."""
    import mmseg_b200  # noqa: F401
    from mmseg_b200 import kernels as K
    from mmseg_b200.src.models.backbones.swin_unetr import SwinUNETR
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf_peak = json.load(open(pk))["bf16_tflops_sustained"] if os.path.exists(pk) else 1400.0
    torch.manual_seed(0)
    m = SwinUNETR(in_channels=2, out_channels=8, feature_size=48).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev)
    res = {"metric": "SwinUNETR-48 forward patches/s (2-channel 96^3, fp16 operands / fp32 accumulate) + bf16 training step",
           "config": "BASELINE.json configs[3] (SwinUNETR feature_size 48, early fusion, 96^3 CT+PET, forward / backward)",
           "peak_tflops": tf_peak,
           "batches": []}
    g = torch.Generator(device="cpu").manual_seed(3)
    for B in (1, 4):
        x = torch.randn((B, 2, 96, 96, 96), generator=g).to(dev)
        with torch.no_grad():
            for _ in range(warmup):
                out = m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = m(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            K.PROFILE = []
            out = m(x)
            torch.cuda.synchronize()
            prof, K.PROFILE = K.PROFILE, None
        kms, flops = {}, 0.0
        for name, info, a, b in prof:
            kms[name] = kms.get(name, 0.0) + a.elapsed_time(b)
            if info and "flops" in info:
                flops += info["flops"]
        row = {"batch": B, "ms": ms, "patches_per_s": B / (ms * 1e-3), "voxels_per_s": B * 96 ** 3 / (ms * 1e-3),
               "gflop_per_patch": flops / B / 1e9, "tflops": flops / (ms * 1e-3) / 1e12,
               "frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / tf_peak, "launches": len(prof),
               "kernel_ms": {k: round(v, 3) for k, v in sorted(kms.items(), key=lambda kv: -kv[1])}}
        res["batches"].append(row)
        log(f"[swin] B={B}: {ms:.2f} ms -> {row['patches_per_s']:.1f} patches/s, {row['tflops']:.0f} TFLOP/s "
            f"({100 * row['frac_of_tensor_peak']:.0f}% of sustained peak), {len(prof)} launches; top: "
            + ", ".join(f"{k[6:]} {v:.2f}" for k, v in list(row["kernel_ms"].items())[:6]))
        if B == 1 and cpu:
            from oracle.swin_unetr import swin_unetr_forward
            torch.set_num_threads(max(1, os.cpu_count() or 1))
            t0 = time.time()
            ref = swin_unetr_forward(sd, x.cpu())
            cpu_s = time.time() - t0
            d = out.cpu() - ref
            res["parity_vs_oracle"] = {"max_abs": d.abs().max().item(), "rel_l2": (d.norm() / ref.norm()).item(),
                                       "label_agreement": (out.cpu().argmax(1) == ref.argmax(1)).double().mean().item(),
                                       "oracle": "oracle/swin_unetr.py (CPU fp32 restatement of MONAI SwinUNETR, PARITY UNPINNED)"}
            res["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"one 96^3 forward through the oracle ({cpu_s:.1f} s)"}
            log(f"[swin] vs oracle: {res['parity_vs_oracle']}; CPU forward {cpu_s:.1f} s")
        del x, out
    # training step (forward + DiceCE + backward + fused AdamW), B = 1 and 2: every op an autograd.Function over the kernels
    try:
        from mmseg_b200.optim import FusedAdamW
        from mmseg_b200.src.trainer.losses import DiceCELoss
        m.train()
        crit = DiceCELoss()
        opt = FusedAdamW(m.parameters(), lr=1e-4)
        res["train"] = []
        for B in (1, 2, 4):
            torch.cuda.reset_peak_memory_stats(dev)     # peak_mem_gb is this batch size's, not the whole process's
            x = torch.randn((B, 2, 96, 96, 96), generator=g).to(dev)
            y = torch.randint(0, 8, (B, 96, 96, 96), generator=g).to(dev)

            def step():
                loss = crit(m(x), y)
                loss.backward()
                opt.step()
                opt.zero_grad()
                return loss

            for _ in range(warmup):
                loss = step()
            torch.cuda.synchronize()
            l0 = K.LAUNCHES[0]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            K.PROFILE = []
            step()
            torch.cuda.synchronize()
            prof, K.PROFILE = K.PROFILE, None
            kms = {}
            for name, info, a, b in prof:
                kms[name] = kms.get(name, 0.0) + a.elapsed_time(b)
            res["train"].append({"batch": B, "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "loss": loss.item(),
                                 "kernel_ms": {k: round(v, 3) for k, v in sorted(kms.items(), key=lambda kv: -kv[1])[:12]},
                                 "kernel_launches_per_step": (K.LAUNCHES[0] - l0) // steps,
                                 "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30})
            log(f"[swin] train B={B}: {ms:.1f} ms/step -> {B / (ms * 1e-3):.1f} samples/s, loss {loss.item():.4f}; top: "
                + ", ".join(f"{k[6:]} {v:.2f}" for k, v in sorted(kms.items(), key=lambda kv: -kv[1])[:8]))
            del x, y
    except Exception as e:   # never allowed to break the line
        torch.cuda.synchronize()
        res["train"] = {"error": f"{type(e).__name__}: {e}"}
    del m
    torch.cuda.empty_cache()
    return res


def run_train(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --workload train needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local)
    sampler.start()
    res = train_measure(args.steps, args.warmup, world, rank, dev, use_graph=not args.no_graph, cpu=not args.no_cpu,
                        which=args.workload)
    sampler.stop_flag = True
    sampler.join()
    if rank == 0:
        res.update({"steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                    "dtype": "bf16", "data": "synthetic", "clocks": sampler.summary()})
        emit(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, library chatter)
    was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto",
                    help="numeric mode of the headline (numerics.py); auto = fastest mode that passes every gate "
                         "(decided from the crop at N=1; the recorded winner at N>1)")
    ap.add_argument("--engine-batch", type=int, default=0, help="windows per engine batch; 0 = chosen per rank")
    ap.add_argument("--no-ladder", action="store_true", help="skip the throughput of the non-headline numeric modes")
    ap.add_argument("--no-selfcheck", action="store_true", help="N>1: skip the sharded-vs-single-GPU label comparison")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity sample")
    ap.add_argument("--workload", default="inference", choices=["inference", "train", "train5", "attention", "swin"],
                    help="inference = headline sliding-window voxels/s (default); train = DualEncoder CT+PET 128^3 B=2 "
                         "samples/s (configs[1]); train5 = 4-modality attention-gate DualEncoder 128^3 B=4 (configs[4])")
    ap.add_argument("--no-graph", action="store_true", help="train workload: do not capture the step in a CUDA graph")
    ap.add_argument("--no-train", action="store_true", help="inference workload: skip the short training measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("train", "train5"):
        run_train(args)
    elif args.workload == "attention":
        torch.cuda.set_device(0)
        emit(attention_measure(torch.device("cuda", 0)))
    elif args.workload == "swin":
        torch.cuda.set_device(0)
        emit(swin_measure(torch.device("cuda", 0), steps=args.steps, warmup=args.warmup, cpu=not args.no_cpu))
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
