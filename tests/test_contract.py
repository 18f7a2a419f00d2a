"""Repository-level contracts that need no GPU: the committed bench line carries every key the bench.py contract names,
and INTEGRATION.md lists every entry point include/mmseg_b200.h declares."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mmseg_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|int32_t|int64_t|const char\*)\s+(mmseg_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def test_committed_bench_lines_follow_the_contract():
    base = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks")
    for name, n in (("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n4.json", 4), ("r01_bench_n8.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for k in base:
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["unit"] == "voxels/s" and d["higher_is_better"] is True and d["scaling"] == "strong"
        assert d["value"] > 0 and d["ms_per_step"] > 0 and d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert "workload" in d["config"] and "model" not in d["config"]
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")))
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and r["frac"] >= 0.5                      # north_star: >= 50 % of the tensor peak on Conv3d
    assert d["roofline_norm"]["bound"] == "hbm" and d["roofline_norm"]["frac"] >= 0.7   # >= 70 % of HBM on the norm kernel
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    p = d["parity"]["parity"]                                    # north_star tolerances, parity mode
    assert p["max_abs"] <= 2e-2 and p["rel_l2"] <= 1e-3 and p["label_agreement"] >= 0.999
    assert abs(p["dice_vs_ref_mean_fg"] - 1.0) <= 1e-3
    # 8-GPU scaling target of north_star (>= 85 %), from the committed lines
    d8 = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n8.json")))
    assert d8["value"] / (8 * d["value"]) >= 0.85


def test_integration_doc_lists_every_entry_point():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    syms = _header_symbols()
    assert len(syms) >= 35
    families = {"mmseg_dicece_bwd": "mmseg_dicece_fwd/_bwd"}     # spelled as a family in the table
    missing = [s for s in syms if s not in doc and families.get(s, "\0") not in doc]
    assert not missing, missing
