"""Compact per-launch summary of an `ncu --page raw --csv` export: python tools/ncu_summary.py raw.csv out.csv [traffic.json note]"""
import csv
import json
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units = rows[hi], rows[hi + 1]
data = [dict(zip(hdr, r)) for r in rows[hi + 2:] if len(r) == len(hdr)]
u = dict(zip(hdr, units))
cols = [c for c in COLS if c in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + [f"{c} [{u[c]}]" for c in cols])
    for d in data:
        w.writerow([d["Kernel Name"].split("(")[0]] + [d[c] for c in cols])
print(f"{len(data)} launches -> {sys.argv[2]}")
if len(sys.argv) > 3:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = sum((float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] +
               float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]) for d in data)
    json.dump({"kernel": "conv3d_tc_kernel + conv3d_roll_kernel", "launches": len(data), "dram_bytes_per_forward_8win": tot,
               "dram_bytes_per_launch_avg": tot / len(data), "note": sys.argv[4] if len(sys.argv) > 4 else "",
               "source": sys.argv[2]}, open(sys.argv[3], "w"), indent=1)
    print(f"traffic {tot / 1e9:.3f} GB per forward -> {sys.argv[3]}")
