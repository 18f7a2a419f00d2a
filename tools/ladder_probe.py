"""Parity / throughput probe of mixed numeric modes: python tools/ladder_probe.py [policy ...]
policy = "A:<bufs>/W:<tags>/R:<tags>" ('+'-joined; "*" = all, "" = none): activation buffers stored hi + lo (in, mid0,
cat0, dec0, pool1, ... bott), conv tags whose weights are split (enc0, enc0.1, dec0, up, ...), conv tags whose raw
output stays fp32.  Prints the four north_star gate metrics on the 36-window bench crop and the full-volume time.
CPU reference = bench.cpu_predictor."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mmseg_b200  # noqa: F401
from mmseg_b200 import _lib
from mmseg_b200.numerics import MODES, NumericMode
from mmseg_b200.src.models.build import build_model
from mmseg_b200.src.trainer.inference import SlidingWindowInferer

policies = sys.argv[1:] or ["A:*/W:enc0/R:*"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_model(bench.model_config("cuda")).eval()
sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
vol_host = bench.synthetic_volume().pin_memory()
predictor, kind, _ = bench.cpu_predictor(sd_cpu)
t, n_win, ref_logits = bench.cpu_sample(predictor, vol_host, bench.CROP, os.cpu_count() or 1)
print(f"cpu {kind}: {t:.1f} s for {n_win} windows", flush=True)
crop_dev = vol_host[:, :bench.CROP[0], :bench.CROP[1], :bench.CROP[2]].contiguous().to(dev)
vol_dev = vol_host.to(dev)
names = []
def _set(spec):
    return None if spec == "*" else frozenset(t for t in spec.split("+") if t)


for pol in policies:
    parts = dict(p.split(":", 1) for p in pol.split("/"))
    A, W, R = _set(parts.get("A", "*")), _set(parts.get("W", "")), _set(parts.get("R", "*"))
    name = f"fp16m[{pol}]"
    MODES[name] = NumericMode(name, _lib.FMT_FP16, A is None, W is None, R is None, A, W, R)
    names.append(name)
extra = os.environ.get("PROBE_EXTRA", "fp16a2,parity").split(",")
for name in names + [e for e in extra if e]:
    model.set_numeric_mode(name)
    ci = SlidingWindowInferer(model, bench.ROI, bench.OVERLAP, bench.MODE, engine_batch=6, use_graph=False)
    got = ci(crop_dev.unsqueeze(0)).cpu()
    m = bench.parity_metrics(got, ref_logits)
    del ci, got
    inf = SlidingWindowInferer(model, bench.ROI, bench.OVERLAP, bench.MODE, engine_batch=8)
    for _ in range(2):
        inf.accumulate(vol_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        inf.accumulate(vol_dev)
        inf.finalize(normalize=False, labels=True)
    e1.record()
    torch.cuda.synchronize()
    m["ms_per_volume"] = e0.elapsed_time(e1) / 2
    print(name, json.dumps(m), flush=True)
    del inf
    model.backbone._engines.pop(name, None)
    torch.cuda.empty_cache()
