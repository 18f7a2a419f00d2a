"""Time one ConvTranspose3d(k2,s2) launch (k1 GEMM + pixel shuffle epilogue):  python tools/convt_dbg.py cin cout S n [flags]
flags & 2 prints the per-CTA cycle counters (needs a scratch stats buffer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmseg_b200
from mmseg_b200 import kernels as K, _lib
from mmseg_b200.kernels import Blocked
cin, cout, S, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
torch.manual_seed(0)
w = torch.randn(cin, cout, 2, 2, 2, device="cuda") * 0.05
b = torch.randn(cout, device="cuda")
pw = K.pack_conv_weight(w, b, False, [cin], transposed=True)
src = Blocked(n, cin, S, S, S, False, "cuda"); src.t.normal_()
dst = Blocked(n, 2 * cout, 2 * S, 2 * S, 2 * S, False, "cuda")
a_cb = K.a_chunk_table(src, [0], [cin], False)
tile = K.plan_conv(S, S, S, n, pw.n_kchunks, pw.n_out, 1, pw.NT)
stats = torch.zeros(n * tile.tiles_per_img * pw.n_out * 2 + 148 * 16, device="cuda") if flags & 2 else None
def run():
    K.conv3d(src, pw, a_cb, dst.t, _lib.OUT_CONVT_K2S2, stats=stats, dst_cbt=dst.cbt, dst_cb_off=0, dst_lo_off=dst.lo_off,
             tile=tile, flags=flags)
for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
us = sorted(ts)[2]
out_mb = n * (2 * S) ** 3 * cout * 2 / 1e6
print(f"flags={flags} convT cin{cin} cout{cout} S{S} n{n} tile {tile.TX,tile.TY,tile.TZ,tile.NT,tile.stages} ntiles {tile.n_ntiles}: {us:.0f} us  "
      f"{out_mb / us * 1e6 / 1e6:.0f} GB/s written")
if flags & 2:
    d = stats.view(torch.int64)[:148 * 8].view(148, 8).double().mean(0).tolist()
    print(f"  clk: mma_total {d[0]:.0f} (a_full wait {d[1]:.0f}, descriptors {d[2]:.0f}, issue {d[3]:.0f}, acc_empty wait {d[4]:.0f}, "
          f"w_full wait {d[7]:.0f})  epi_total {d[5]:.0f}  epi_wait_acc_full {d[6]:.0f}")
