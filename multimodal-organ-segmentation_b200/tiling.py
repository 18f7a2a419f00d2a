"""Host-side tile planner for the tcgen05 conv kernel (csrc/conv_tc.cu).

Pure integer arithmetic — mirrors plan_conv() in conv_tc.cu so that a plan chosen here is accepted there — plus a
small cycle model used to pick (TX, TY, TZ, NT, stages) per layer.  Hardware constants are B200's: 148 SMs,
227 KB shared memory per SM, 512 TMEM columns per SM, 4096 bf16 MAC/clk/SM; measured on B200 (tools/micro/umma_bench.cu):
an M=128, K=16 tcgen05.mma costs max(51, N/2) clk, i.e. the tensor pipe is only saturated for N >= ~104, which is why
the kernel folds the three z taps into N (N = 3*NT for interior planes).
"""
import os
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional

NUM_SMS = 148
SMEM_LIMIT = 227 * 1024
SMEM_HALF = 113 * 1024       # two CTAs per SM
HEADER_BYTES = 2048
TMEM_COLS = 512
MAX_MT = 8
MMA_MIN_CYCLES = 51.0
L2_BYTES_PER_CLK_SM = 36.0   # L2 -> SM when every SM is loading (~6300 B/clk chip-wide)


def _round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def _cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


@dataclass(frozen=True)
class ConvTile:
    TX: int
    TY: int
    TZ: int
    NT: int
    n_ntiles: int
    stages: int
    mt: int
    smem_bytes: int
    tiles_per_img: int
    est_cycles: float
    roll: bool = False       # rolling-z kernel: TZ is the z-segment length, tiles_per_img counts (x, y, z-segment) items
    kpb: int = 1             # rolling-z: K chunks per TMA stage (2 = paired, needs adjacent channel blocks)


def tmem_cols(mt: int, TZ: int, NT: int) -> int:
    """TMEM allocation: two accumulator sets when they fit in 512 columns (epilogue / MMA overlap), else one."""
    cols, tc = mt * TZ * NT, 32
    if 2 * cols <= TMEM_COLS:
        cols *= 2
    while tc < cols:
        tc <<= 1
    return tc


def w_stages_for(ksize: int, NT: int, n_kchunks: int) -> int:
    """Weight slots: 2 for k=3; k=1 keeps the whole [K x NT] panel resident when it fits 64 KB (mirrors plan_conv)."""
    if ksize != 1:
        return 2
    resident = n_kchunks * _round_up(NT * 32, 128) <= 64 * 1024
    return n_kchunks if (resident and n_kchunks > 8) else 8


def smem_bytes(ksize: int, TX: int, TY: int, NT: int, stages: int, TZ: int = 1, n_kchunks: int = 1) -> Optional[int]:
    """Dynamic shared memory of one CTA, or None if the tiling is invalid (same arithmetic as plan_conv)."""
    h = ksize // 2
    PX, PY = TX + 2 * h, TY + 2 * h
    if PX > 128 or PY > 256:
        return None
    flat = ksize == 1                      # k=1: the whole TX x TY x TZ tile is one stage / one set of M tiles
    mt = _cdiv(TX * TY * TZ, 128) if flat else _cdiv((TY - 1) * PX + TX, 128)
    n_acc = mt if flat else mt * TZ
    if mt > MAX_MT or n_acc * NT > TMEM_COLS or (ksize == 3 and 3 * NT > 256) or (flat and TZ > 256):
        return None
    taps = ksize ** 3
    w_bytes = taps * NT * 32
    plane = PX * PY * 16 * (TZ if flat else 1)
    if (plane >> 4) > 0x3FFF:
        return None
    a_tx = 2 * plane
    stage = _round_up(a_tx, 128)
    rows_needed = mt * 128 + 2 * h * PX + 2 * h
    overflow = max(rows_needed * 16 - plane, 0)
    w_stages = w_stages_for(ksize, NT, n_kchunks)
    total = HEADER_BYTES + w_stages * _round_up(w_bytes, 128) + stages * stage + _round_up(overflow, 128) + 128
    if tmem_cols(mt, 1 if flat else TZ, NT) > 256 and total < 116 * 1024:
        total = 116 * 1024
    if total > SMEM_LIMIT or w_bytes >= (1 << 20) or a_tx >= (1 << 20):
        return None
    return total


def _mma_cycles(n: int) -> float:
    return max(MMA_MIN_CYCLES, n / 2.0)


@lru_cache(maxsize=None)
def plan_conv(X: int, Y: int, Z: int, n_img: int, n_kchunks: int, n_out: int, ksize: int,
              nt_cap: int = 64, epi_cost: float = 150.0) -> ConvTile:
    """Pick a tiling for a conv with n_out GEMM columns (already padded to a multiple of 16)."""
    h = ksize // 2
    NT = min(n_out, nt_cap)
    while n_out % NT:
        NT -= 16
    n_ntiles = n_out // NT
    best = None
    nx_min = 1
    while _cdiv(X, nx_min) + 2 * h > 128:
        nx_min += 1
    for nx in range(nx_min, min(nx_min + 3, X) + 1):
        TX = _cdiv(X, nx)
        PX = TX + 2 * h
        for TY in range(1, Y + 1):
            mt = _cdiv((TY - 1) * PX + TX, 128)
            if not h:
                mt = _cdiv(TX * TY, 128)
            if mt > MAX_MT or mt * NT > TMEM_COLS:
                break
            tz_max = min(Z, TMEM_COLS // (mt * NT)) if h else min(Z, 256, (MAX_MT * 128) // (TX * TY))
            for TZ in range(1, tz_max + 1):
                if not h:
                    mt = _cdiv(TX * TY * TZ, 128)
                    if mt * NT > TMEM_COLS:
                        break
                    if mt in (3, 5, 7) and os.environ.get("MMSEG_K1_ODD_MT", "0") != "1":
                        continue   # the flat epilogue's two-chunk pipeline spills 340 B for odd M-tile counts (ptxas log)
                tc = tmem_cols(mt, TZ if h else 1, NT)
                sb, stages = None, 0
                # deep ring: one stage is a single small z-plane (a few KB) while a TMA round trip is ~1.5-2 us (a third
                # of the activation bytes miss L2), so the bytes in flight — ring depth x stage size — bound the load
                # rate (measured: 8 x 5.8 KB in flight = 20 GB/s per SM = load-bound).  Take the deepest ring that fits.
                for st in (24, 20, 16, 12, 8, 6, 4, 3, 2):
                    s_ = smem_bytes(ksize, TX, TY, NT, st, TZ, n_kchunks)
                    if s_ is not None:
                        sb, stages = s_, st
                        break
                if sb is None:
                    continue
                tx_n, ty_n, tz_n = _cdiv(X, TX), _cdiv(Y, TY), _cdiv(Z, TZ)
                n_tiles = tx_n * ty_n * tz_n * n_img
                if h:
                    per_tap = sum(_mma_cycles(NT * (min(2, pl) - max(0, pl - TZ + 1) + 1)) for pl in range(TZ + 2))
                    mma = n_kchunks * 9 * mt * (per_tap + 0.0)
                else:   # one tcgen05.commit (~45 clk) per stage = per K chunk, plus the TMA issue rate (~300 clk / stage)
                    mma = n_kchunks * max(mt * _mma_cycles(NT) + 45.0, 300.0)
                planes = TZ + 2 * h
                load = n_kchunks * (planes * 2 * PX * (TY + 2 * h) * 16 + ksize ** 3 * NT * 32) / L2_BYTES_PER_CLK_SM
                epi = (TZ if h else 1) * mt * (NT // 16) * epi_cost + 500.0
                # persistent CTAs (about one per SM in total): with two accumulator sets the epilogue of a tile hides
                # behind the MMAs of the next one
                double = 2 * mt * (TZ if h else 1) * NT <= TMEM_COLS
                per_tile = max(mma, load, epi) + 300.0 if double else max(mma, load) + epi
                ctas = max(1, min(n_tiles, NUM_SMS // n_ntiles))
                est = _cdiv(n_tiles, ctas) * per_tile + 4000.0
                cand = (est, -TY * TZ * TX)
                if best is None or cand < best[0]:
                    best = (cand, ConvTile(TX, TY, TZ, NT, n_ntiles, stages, mt, sb, tx_n * ty_n * tz_n, est))
    if best is None:
        raise ValueError(f"no valid conv tiling for X={X} Y={Y} Z={Z} n_out={n_out} ksize={ksize}")
    return best[1]


ROLL_FLAG = 16               # MMSEG_CONV_ROLL_Z
ROLL_KPAIR_FLAG = 32         # MMSEG_CONV_ROLL_KPAIR


def roll_smem_bytes(TX: int, TY: int, NT: int, n_kchunks: int, stages: int, kpb: int = 1) -> Optional[int]:
    """Shared memory of conv3d_roll_kernel (all K-chunk weights resident), or None when the tiling is invalid."""
    PX, PY = TX + 2, TY + 2
    if PX > 128 or PY > 256 or _cdiv((TY - 1) * PX + TX, 128) != 1 or NT != 32:
        return None
    plane = PX * PY * 16
    stage = _round_up(2 * plane * kpb, 128)
    overflow = max((128 + 2 * PX + 2) * 16 - plane, 0)
    total = HEADER_BYTES + n_kchunks * _round_up(27 * NT * 32, 128) + stages * stage + _round_up(overflow, 128) + 128
    return total if total <= SMEM_LIMIT else None


@lru_cache(maxsize=None)
def plan_roll(X: int, Y: int, Z: int, n_img: int, n_kchunks: int, n_out: int, kpb: int = 1) -> Optional[ConvTile]:
    """Rolling-z plan for a k=3, C_out = 32 layer (None when not eligible).  Cost model: an N = 96 MMA is
    shared-memory-bound at ~56 clk (measured), a partial-N boundary MMA ~51; one item = one z segment of a column."""
    if n_out != 32:
        return None
    best = None
    nx_min = 1
    while _cdiv(X, nx_min) + 2 > 128:
        nx_min += 1
    for nx in range(nx_min, min(nx_min + 6, X) + 1):
        TX = _cdiv(X, nx)
        PX = TX + 2
        TY = 0
        while TY < Y and (TY * PX + TX) <= 128:   # largest TY with (TY-1)*PX + TX <= 128
            TY += 1
        if TY < 1:
            continue
        stages = 0
        for st in (24, 20, 16, 12, 10, 8, 6, 5, 4):
            sb = roll_smem_bytes(TX, TY, 32, n_kchunks, st, kpb)
            if sb is not None:
                stages = st
                break
        if stages * kpb < max(8, 2 * n_kchunks):
            continue
        tx_n, ty_n = _cdiv(X, TX), _cdiv(Y, TY)
        for nseg in range(1, max(1, Z // 8) + 1):
            ZS = _cdiv(Z, nseg)
            nseg_eff = _cdiv(Z, ZS)
            items = tx_n * ty_n * nseg_eff * n_img
            per_item = n_kchunks * 9 * ((ZS - 2) * 56.0 + 4 * 51.0) + ZS * 9 * 3.5 * n_kchunks + 1500.0
            est = _cdiv(items, NUM_SMS) * per_item + 4000.0
            cand = (est, -TX * TY)
            if best is None or cand < best[0]:
                best = (cand, ConvTile(TX, TY, ZS, 32, 1, stages, 1, sb, tx_n * ty_n * nseg_eff, est, True, kpb))
    return None if best is None else best[1]
