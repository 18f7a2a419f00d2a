"""Stand-alone GPU case runner: one subprocess (with a timeout) per case so that a trapped or hung kernel is reported
and cannot poison the CUDA context of the next case.  Usage on the GPU box:
    python tools/gpu_check.py [group ...] > gpurun_out/check.log
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "conv0": [
        "conv_case(16, 32, (4, 6, 20), name='first')",
    ],
    "conv": [
        "conv_case(16, 32, (8, 12, 20))",
        "conv_case(2, 32, (6, 10, 24))",
        "conv_case(32, 32, (7, 9, 30), n_img=2, tile=(30, 4, 2))",
        "conv_case(32, 64, (5, 11, 50), tile=(25, 5, 2))",
        "conv_case(64, 128, (6, 6, 6), n_img=2)",
        "conv_case(128, 256, (6, 6, 6))",
        "conv_case(32, 32, (16, 24, 96))",
        "conv_case(32, 32, (6, 8, 20), split=True)",
        "conv_case(64, 32, (8, 16, 48), split=True)",
        "conv_case(32, 16, (5, 6, 20), ks=1)",
    ],
    "block": [
        "conv_block_case(16, 32, (8, 12, 16))",
        "conv_block_case(2, 32, (8, 12, 16), split=True)",
        "conv_block_case(32, 64, (8, 8, 16), n_img=2, slope=0.2, split=True)",
        "convt_case(64, (3, 5, 6))",
        "convt_case(32, (4, 4, 12), n_img=2, split=True)",
        "logits_case(32, 8, (6, 7, 20))",
        "logits_case(32, 8, (6, 7, 20), split=True)",
        "pack_roundtrip_case()",
    ],
    "unet": [
        "unet_case((16, 32, 64), 32, 1, 'parity')",
        "unet_case((16, 32, 64), 32, 2, 'bf16')",
        "unet_case((32, 64, 128, 256, 512), 96, 1, 'parity')",
        "unet_case((32, 64, 128, 256, 512), 96, 1, 'bf16')",
    ],
    "time": [
        "unet_time_case(1, 96, 'bf16')",
        "unet_time_case(4, 96, 'bf16')",
        "unet_time_case(1, 96, 'parity')",
    ],
}


def run_case(expr: str, timeout: int = 180) -> bool:
    code = f"import sys; sys.path.insert(0, {ROOT!r}); from tests.gpu_cases import *; {expr}"
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
        ok = p.returncode == 0
        out = (p.stdout + p.stderr).strip().splitlines()
    except subprocess.TimeoutExpired as e:
        ok, out = False, [f"TIMEOUT after {timeout}s", *(str(e.stdout or "").splitlines()[-5:])]
    tail = out if ok else out[-25:]
    print(f"{'PASS' if ok else 'FAIL'} {expr}  ({time.time() - t0:.1f}s)")
    for line in tail:
        print("    " + line)
    sys.stdout.flush()
    return ok


def main():
    groups = sys.argv[1:] or list(CASES)
    n_fail = 0
    for g in groups:
        print(f"==== {g}")
        for expr in CASES[g]:
            n_fail += 0 if run_case(expr) else 1
    print(f"==== done, {n_fail} failed")
    return 1 if n_fail else 0


if __name__ == "__main__":
    sys.exit(main())
