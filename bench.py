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


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_predictor(sd):
    from oracle.models import unet3d_forward
    return lambda w: unet3d_forward(sd, w)


def cpu_sample(sd, vol_cpu, crop, threads):
    """Oracle sliding window (reference algorithm, CPU fp32) over a crop; returns (seconds, n_windows, logits)."""
    from oracle.sliding_window import sliding_window_inference as oswi, window_starts
    torch.set_num_threads(threads)
    x = vol_cpu[:, :crop[0], :crop[1], :crop[2]].unsqueeze(0).contiguous()
    n_win = len(window_starts(crop, ROI, OVERLAP))
    t0 = time.perf_counter()
    with torch.no_grad():
        out = oswi(x, ROI, 4, oracle_predictor(sd), overlap=OVERLAP, mode=MODE)
    return time.perf_counter() - t0, n_win, out


def full_volume_equiv(seconds, n_win):
    """voxels/s of the full 512x512x300 job extrapolated from `n_win` windows (600 windows per volume)."""
    from oracle.sliding_window import window_starts
    total = len(window_starts(VOL, ROI, OVERLAP))
    return (VOL[0] * VOL[1] * VOL[2]) / (seconds / n_win * total)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference is PyTorch and cannot travel to the GPU
    box; MONAI, which it delegates the sliding window to, is not installed anywhere).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sys.path.insert(0, ROOT)
    import mmseg_b200  # noqa: F401  (only for the module tree that creates the same random-init weights)
    from mmseg_b200.src.models.backbones.unet import UNet3D
    m = UNet3D(2, 8, FEATURES).eval()
    sd = {"backbone." + k: v for k, v in m.state_dict().items()}
    crop = (96, 144, 144)  # 1 x 2 x 2 = 4 windows = one sw_batch of 4 (trainer.py:386-392 uses sw_batch_size=4)
    vol = synthetic_volume(crop, 1234)
    times = []
    for i in range(args.warmup + args.steps):
        t, n_win, _ = cpu_sample(sd, vol, crop, threads)
        if i >= args.warmup:
            times.append(t)
        log(f"[reference] step {i}: {t:.2f} s for {n_win} windows")
    ms = sum(times) / len(times) * 1e3
    v = full_volume_equiv(ms / 1e3, 4)
    sample = "4 of the 600 windows per step (96x144x144 crop, sw_batch 4), extrapolated x150 to the full volume"
    emit({
        "impl": "reference", "metric": "sliding-window inference voxels/s (CT+PET)", "value": v, "unit": "voxels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, world):
    return {"workload": "sliding_window_inference 2ch CT+PET 512x512x300, UNet3D early fusion features 32-512, "
                        "roi 96^3 overlap 0.5 gaussian blend, 600 windows -> uint8 labels (BASELINE.json configs[2])",
            "windows": 600, "engine_batch": args.engine_batch, "numeric_mode": args.mode,
            "parallelism": f"window-chunks x{world} (axis-0 slabs, one partial-sum exchange)" if world > 1 else "single GPU",
            "l2": "inputs_exceed_l2 (volume 629 MB + accumulator 2.8 GB per step >> 126 MB L2)"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a B200; there is no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import mmseg_b200  # noqa: F401
    from mmseg_b200 import kernels as K
    from mmseg_b200.src.models.build import build_model
    from mmseg_b200.src.trainer.inference import SlidingWindowInferer, predict_volume, shard_windows, _INFERERS

    torch.manual_seed(0)
    model = build_model(model_config("cuda")).eval()
    model.set_numeric_mode(args.mode)
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}

    t0 = time.time()
    vol_host = synthetic_volume().pin_memory()
    log(f"[rank {rank}] synthetic volume {tuple(vol_host.shape)} in {time.time() - t0:.1f} s")
    inf = SlidingWindowInferer(model, ROI, OVERLAP, MODE, engine_batch=args.engine_batch)
    _INFERERS[(id(model), ROI, OVERLAP, MODE, args.engine_batch)] = inf   # predict_volume() reuses this engine
    vol_dev = inf.device_volume(vol_host.shape, dev)
    vol_dev.copy_(vol_host)
    nvox = VOL[0] * VOL[1] * VOL[2]
    n_windows = 600
    lo, hi = shard_windows(n_windows, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world > 1:
            return inf.run_sharded(vol_dev)
        inf.accumulate(vol_dev)
        return inf.finalize(normalize=False, labels=True)[1]

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

    sampler = ClockSampler(local)
    sampler.start()
    ms_step, launches, labels = timed(step_resident, args.steps, args.warmup)
    sampler.stop_flag = True
    sampler.join()
    value = nvox / (ms_step * 1e-3)
    log(f"[rank {rank}] resident: {ms_step:.2f} ms/volume -> {value / 1e6:.1f} Mvox/s")

    # ---- end to end through the public host API (pinned host in, pinned host out)
    out_host = torch.empty(VOL, dtype=torch.uint8).pin_memory()
    z0, z1 = inf.input_range(VOL, world, rank)
    h2d = torch.tensor([2 * (z1 - z0) * VOL[1] * VOL[2] * 4], device=dev, dtype=torch.int64)
    d2h = torch.tensor([nvox], device=dev, dtype=torch.int64)   # every rank reads the gathered label map back
    if world > 1:
        dist.all_reduce(h2d)
        dist.all_reduce(d2h)
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e, _, _ = timed(lambda: predict_volume(model, vol_host, ROI, OVERLAP, MODE, args.engine_batch,
                                                out_host=out_host), e2e_steps, 1)
    assert torch.equal(out_host, labels.cpu()), "end-to-end labels differ from the device-resident run"
    e2e = {"value": nvox / (ms_e2e * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": int(h2d.item()),
           "d2h_bytes_per_step": int(d2h.item()), "ms_per_step": ms_e2e, "steps": e2e_steps}
    log(f"[rank {rank}] e2e: {ms_e2e:.2f} ms/volume -> {e2e['value'] / 1e6:.1f} Mvox/s")

    def secondary_train():
        """training samples/s (BASELINE.json configs[1]); never allowed to break the headline line"""
        if args.no_train:
            return None
        try:
            inf._state = None
            inf._dev_vol = None
            _INFERERS.clear()
            torch.cuda.empty_cache()
            return train_measure(3, 2, world, rank, dev, use_graph=True, cpu=False)
        except Exception as e:
            torch.cuda.synchronize()
            return {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        secondary_train()
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
    st = inf._state
    nb = st["nb"]
    st["slots"][0]["starts_dev"][:nb].copy_(st["starts_all"][:nb])
    inf._run_batch(st, vol_dev, nb)
    torch.cuda.synchronize()
    K.PROFILE = []
    for _ in range(3):
        inf._run_batch(st, vol_dev, nb)
    torch.cuda.synchronize()
    prof, K.PROFILE = K.PROFILE, None
    agg, layers = {}, {}
    for name, info, a, b in prof:
        ms = a.elapsed_time(b)
        d = agg.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        d["ms"] += ms
        d["n"] += 1
        if info:
            d["flops"] += info.get("flops", 0.0)
            d["bytes"] += info.get("bytes", 0.0)
            L = layers.setdefault((name, info["layer"]), {"ms": 0.0, "n": 0, "info": info})
            L["ms"] += ms
            L["n"] += 1
    total_ms = sum(d["ms"] for d in agg.values())
    conv, norm = agg["mmseg_conv3d_fwd"], agg["mmseg_instnorm_act_apply"]
    conv_tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
    norm_gbs = norm["bytes"] / (norm["ms"] * 1e-3) / 1e9
    log(f"per-kernel shares of one {nb}-window batch (CUDA events, 3 repeats):")
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
    traffic = None   # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture
    tj = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.exists(tj):
        traffic = json.load(open(tj)).get("dram_bytes_per_launch_avg")
    roofline = {"bound": "tensor", "kernel": "conv3d_tc_kernel + conv3d_roll_kernel", "achieved": conv_tf, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": conv_tf / tf_peak, "frac_of_nominal_2250": conv_tf / 2250.0, "traffic": traffic,
                "traffic_note": "average DRAM bytes per conv launch (ncu --set full, same 8-window forward; profiles/r01_v5_ncu_conv_launches.csv)",
                "peak_source": peak_src,
                "launches": conv["n"] // 3, "avg_launch_ms": conv["ms"] / conv["n"],
                "share_of_step": conv["ms"] / total_ms,
                "note": "algorithmic conv FLOPs of one %d-window batch / sum of conv launch durations" % nb}
    roofline_norm = {"bound": "hbm", "kernel": "instnorm_apply(_pool)_kernel", "achieved": norm_gbs, "peak": hbm_peak,
                     "unit": "GB/s", "frac": norm_gbs / hbm_peak, "frac_of_nominal_8000": norm_gbs / 8000.0, "traffic": None,
                     "share_of_step": norm["ms"] / total_ms, "avg_launch_ms": norm["ms"] / norm["n"]}
    step_tf = n_windows * GF_PER_WINDOW / 1e3 / (ms_step * 1e-3) / world
    # ---- CPU baseline + parity on the bounded sample (rank 0, N=1 only)
    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sd_o = {k: v for k, v in sd_cpu.items()}
        vol_cpu = vol_host.clone()
        t, n_win, ref_logits = cpu_sample(sd_o, vol_cpu, CROP, threads)
        cpu_baseline = {"value": full_volume_equiv(t, n_win), "unit": "voxels/s", "cores": threads, "kind": "port",
                        "sample": f"{n_win} of the 600 windows ({CROP[0]}x{CROP[1]}x{CROP[2]} crop of the same volume, "
                                  f"{t:.1f} s), extrapolated x600/{n_win}; oracle = CPU fp32 restatement of the reference"}
        ref_lab = ref_logits.argmax(1)[0]
        parity = {}
        crop_dev = vol_dev[:, :CROP[0], :CROP[1], :CROP[2]].contiguous()
        for mode in ("parity", "bf16"):
            model.set_numeric_mode(mode)
            ci = SlidingWindowInferer(model, ROI, OVERLAP, MODE, engine_batch=6, use_graph=False)
            got = ci(crop_dev.unsqueeze(0)).cpu()
            lab = got.argmax(1)[0]
            dices = []
            for c in range(1, 8):
                a, b = (lab == c), (ref_lab == c)
                dices.append((2.0 * (a & b).sum().item() + 1e-5) / (a.sum().item() + b.sum().item() + 1e-5))
            d = (got - ref_logits).double()
            parity[mode] = {"max_abs": d.abs().max().item(), "rel_l2": (d.norm() / ref_logits.double().norm()).item(),
                            "label_agreement": (lab == ref_lab).double().mean().item(),
                            "dice_vs_ref_mean_fg": sum(dices) / len(dices)}
        model.set_numeric_mode(args.mode)
        log("parity vs oracle on the crop:", json.dumps(parity))

    vol_dev = None
    train_res = secondary_train()

    line = {
        "metric": "sliding-window inference voxels/s (CT+PET)", "value": value, "unit": "voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "bf16x3",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
        "roofline": roofline, "roofline_norm": roofline_norm,
        "step_tflops_per_gpu": step_tf, "step_frac_of_tensor_peak": step_tf / tf_peak,
        "cpu_baseline": cpu_baseline, "parity": parity, "train": train_res,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------ training workload
TRAIN_S, TRAIN_B = 128, 2
TRAIN_GF_PER_STEP = 7394.0 * (TRAIN_B / 2)   # DualEncoder C2 train step, per GPU


def train_config(device):
    return {"model": {"name": "dual_encoder", "in_channels": 2, "out_channels": 8,
                      "backbone": {"features": FEATURES, "norm": "instance"},
                      "fusion": {"type": "cross_attention"}, "head": {"dropout": 0.1}},
            "data": {"modalities": ["CT", "PET"]},
            "training": {"epochs": 1, "optimizer": {"name": "adamw", "lr": 1e-4, "weight_decay": 1e-5},
                         "loss": {"name": "dice_ce", "dice_weight": 0.5, "ce_weight": 0.5}, "accumulation_steps": 1},
            "hardware": {"device": device, "mixed_precision": True, "cuda_graph": True},
            "experiment": {"output_dir": "/tmp/mmseg_b200_bench", "name": "train"}}


def train_measure(steps, warmup, world, rank, dev, use_graph=True, cpu=False):
    """BASELINE.json configs[1]: DualEncoder (fusion 'cross_attention' == mean over modalities in the reference) on
    CT+PET 128^3 patches, batch 2 per GPU, bf16 kernels, DiceCE, AdamW; data-parallel over `world` ranks (weak scaling).
    Returns a dict with device-resident and end-to-end (pinned host batch in, loss out) samples/s."""
    import torch.distributed as dist
    import mmseg_b200  # noqa: F401
    from mmseg_b200 import kernels as K
    from mmseg_b200.src.models.build import build_model
    from mmseg_b200.src.trainer.trainer import Trainer
    torch.manual_seed(0)                       # identical replicas
    cfg = train_config("cuda")
    model = build_model(cfg)
    tr = Trainer(cfg, model)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand((TRAIN_B, 2, TRAIN_S, TRAIN_S, TRAIN_S), generator=g).pin_memory()
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

    l0 = K.LAUNCHES[0]
    ms, loss = timed(lambda: step(x, y), steps, warmup)
    launches = (K.LAUNCHES[0] - l0) if mode == "eager" else None

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        return step(xd, yd).item()          # device->host read of the loss, like trainer.py:260
    ms_e2e, last = timed(e2e_step, max(1, min(steps, 5)), 1)
    res = {"metric": "training samples/s (DualEncoder CT+PET 128^3, batch 2/GPU, bf16 step)",
           "value": world * TRAIN_B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "n_gpus": world,
           "scaling": "weak", "mode": mode, "loss": float(loss.item() if torch.is_tensor(loss) else loss),
           "e2e": {"value": world * TRAIN_B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": world * (x_host.numel() * 4 + y_host.numel() * 8), "d2h_bytes_per_step": world * 4},
           "tflops_per_gpu": TRAIN_GF_PER_STEP / 1e3 / (ms * 1e-3), "gpu_launches_per_step": None,
           "config": {"workload": "DualEncoder(fusion=cross_attention -> mean) 2 modalities 128^3, batch 2/GPU, "
                                  "features 32-512, dropout 0.1, DiceCE, AdamW, fwd+bwd+allreduce+step "
                                  "(BASELINE.json configs[1])", "parallelism": f"dp{world}"}}
    # one eager profiled step: per-kernel shares (also counts the launches of a step)
    K.PROFILE = []
    l0 = K.LAUNCHES[0]
    tr.train_step(x, y)
    torch.cuda.synchronize()
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
    res["kernel_tflops"] = {k: round(v[1] / v[0] / 1e9, 1) for k, v in agg.items() if v[1] > 0}
    res["kernel_gbs"] = {k: round(v[2] / v[0] / 1e6, 0) for k, v in agg.items() if v[2] > 0}
    if rank == 0:
        log(f"train step {mode}: {ms:.2f} ms -> {res['value']:.1f} samples/s ({res['tflops_per_gpu']:.0f} TFLOP/s/GPU); "
            f"e2e {ms_e2e:.2f} ms; kernels {tot:.2f} ms: " + ", ".join(f"{k.replace('mmseg_', '')} {v:.2f}" for k, v in list(res['kernel_ms'].items())[:8]))
    if cpu and rank == 0 and world == 1:
        from oracle.train import train_step as oracle_step
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k[len("backbone."):]: v.detach().cpu() for k, v in tr.model.state_dict().items()}
        t0 = time.perf_counter()
        oracle_step("dual", sd, dict(L=len(FEATURES), M=2, fusion="cross_attention"), x_host[:1], y_host[:1], torch.float32)
        t = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 1.0 / t, "unit": "samples/s", "cores": threads, "kind": "port",
                               "sample": f"one fwd+DiceCE+bwd step at batch 1 ({t:.1f} s) through the oracle (CPU fp32 autograd)"}
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
    res = train_measure(args.steps, args.warmup, world, rank, dev, use_graph=not args.no_graph, cpu=not args.no_cpu)
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
    ap.add_argument("--mode", default="bf16", choices=["bf16", "parity"])
    ap.add_argument("--engine-batch", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity sample")
    ap.add_argument("--workload", default="inference", choices=["inference", "train"],
                    help="inference = headline sliding-window voxels/s (default); train = DualEncoder 128^3 samples/s")
    ap.add_argument("--no-graph", action="store_true", help="train workload: do not capture the step in a CUDA graph")
    ap.add_argument("--no-train", action="store_true", help="inference workload: skip the short training measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
