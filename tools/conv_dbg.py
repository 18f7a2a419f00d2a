"""Time single conv layers (persistent tcgen05 kernel):  python tools/conv_dbg.py cin cout S n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmseg_b200
from mmseg_b200 import kernels as K, _lib
from mmseg_b200.kernels import Blocked
cin, cout, S, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torch.manual_seed(0)
w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
pw = K.pack_conv_weight(w, None, False, [cin], use_bias=False)
src = Blocked(n, (cin + 15) // 16 * 16, S, S, S, False, "cuda"); src.t.normal_()
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = torch.empty((n, cout // 8, S, S, S, 8), dtype=torch.bfloat16, device="cuda")
a_cb = K.a_chunk_table(src, [0], [cin], False)
tile = K.plan_conv_norm((S, S, S), n, pw, False, a_cb)
stats = torch.zeros(n * tile.tiles_per_img * cout * 2, device="cuda")
for _ in range(3):
    K.conv3d(src, pw, a_cb, raw, _lib.OUT_BLOCKED_BF16, stats=stats, dst_cbt=cout // 8, tile=tile, flags=flags)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.conv3d(src, pw, a_cb, raw, _lib.OUT_BLOCKED_BF16, stats=stats, dst_cbt=cout // 8, tile=tile, flags=flags)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
us = sorted(ts)[2]
print(("roll " if tile.roll else "") + f"flags={flags} cin{cin} cout{cout} S{S} n{n} tile {tile.TX,tile.TY,tile.TZ,tile.NT,tile.stages}: {us:.0f} us  {2.0*n*S**3*cin*cout*27/us/1e6:.0f} TF/s")
if flags & 2:   # per-CTA counters written over the stats buffer: [0] MMA warp total, [5] epilogue total, [6] epilogue waiting
    torch.cuda.synchronize()
    d = stats.view(torch.int64)[:148 * 8].view(148, 8).double().mean(0).tolist()
    if tile.roll:
        print(f"  roll: mma_total {d[0]:.0f} clk, stages {d[3]:.0f} ({d[0] / max(d[3], 1):.0f} clk/stage), a_full not ready {d[1]:.0f}, "
              f"z_empty not ready {d[2]:.0f}; epi_total {d[5]:.0f}, epi waiting z_full {d[6]:.0f}")
    print(f"  clk: mma_total {d[0]:.0f} (a_full wait {d[1]:.0f}, descriptors {d[2]:.0f}, issue {d[3]:.0f}, acc_empty wait {d[4]:.0f}, w_full wait {d[7]:.0f})  epi_total {d[5]:.0f}  epi_wait_acc_full {d[6]:.0f}")
