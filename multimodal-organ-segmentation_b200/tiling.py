"""Host-side tile planner for the tcgen05 conv kernel (csrc/conv_tc.cu).

Pure integer arithmetic — mirrors plan_conv() in conv_tc.cu so that a plan chosen here is accepted there — plus a
small cycle model used to pick (TX, TY, TZ, NT, stages) per layer.  Hardware constants are B200's: 148 SMs,
227 KB shared memory per CTA, 512 TMEM columns, 4096 bf16 MAC/clk/SM, 128 B/clk shared-memory read, and
~40 B/clk/SM of L2->SM bandwidth when every SM is loading.
"""
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional

NUM_SMS = 148
SMEM_LIMIT = 227 * 1024
HEADER_BYTES = 1024
TMEM_COLS = 512
MAX_ACC = 32


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


def smem_bytes(ksize: int, TX: int, TY: int, NT: int, stages: int) -> Optional[int]:
    """Dynamic shared memory of one CTA, or None if the tiling is invalid (same arithmetic as plan_conv)."""
    h = ksize // 2
    PX, PY = TX + 2 * h, TY + 2 * h
    if PX > 128 or PY > 256:
        return None
    mt = _cdiv((TY - 1) * PX + TX, 128)
    taps = ksize ** 3
    w_bytes = taps * NT * 32
    plane = PX * PY * 16
    if (plane >> 4) > 0x3FFF:
        return None
    a_tx = 2 * plane
    stage = _round_up(a_tx, 128)
    rows_needed = mt * 128 + 2 * h * PX + 2 * h
    overflow = max(rows_needed * 16 - plane, 0)
    total = HEADER_BYTES + 2 * _round_up(w_bytes, 128) + stages * stage + _round_up(overflow, 128) + 128
    if total > SMEM_LIMIT or w_bytes >= (1 << 20) or a_tx >= (1 << 20):
        return None
    return total


def _mma_cycles(NT: int) -> float:
    # one UMMA 128 x NT x 16: tensor floor NT/2 clk; shared-memory operand read (4 KB of A + NT*32 B of B) at 128 B/clk
    return max(NT / 2.0, (4096 + NT * 32) / 128.0)


@lru_cache(maxsize=None)
def plan_conv(X: int, Y: int, Z: int, n_img: int, n_kchunks: int, n_out: int, ksize: int,
              nt_cap: int = 64) -> ConvTile:
    """Pick a tiling for a conv with n_out GEMM columns (already padded to a multiple of 16)."""
    h = ksize // 2
    NT = min(n_out, nt_cap)
    while n_out % NT:
        NT -= 16
    n_ntiles = n_out // NT
    max_acc = min(MAX_ACC, TMEM_COLS // NT)
    # X tiling: whole rows when the TMA box allows it, otherwise equal parts
    nx = 1
    while _cdiv(X, nx) + 2 * h > 128:
        nx += 1
    TX = _cdiv(X, nx)
    PX = TX + 2 * h
    best = None
    for TY in range(1, Y + 1):
        mt = _cdiv((TY - 1) * PX + TX, 128)
        if mt > max_acc:
            break
        for TZ in range(1, min(Z, max_acc // mt) + 1):
            for stages in (4, 3, 2):
                sb = smem_bytes(ksize, TX, TY, NT, stages)
                if sb is not None:
                    break
            else:
                continue
            ty, tz = _cdiv(Y, TY), _cdiv(Z, TZ)
            n_cta = nx * ty * tz * n_img * n_ntiles
            # per-CTA work for a full interior tile
            pairs = TZ * (3 if h else 1)            # (input plane, dz) pairs that issue MMAs
            mma = n_kchunks * pairs * (9 if h else 1) * mt * _mma_cycles(NT)
            planes = TZ + 2 * h
            load = n_kchunks * (planes * 2 * PX * (TY + 2 * h) * 16 + ksize ** 3 * NT * 32) / 40.0
            epi = TZ * mt * NT * 8.0 + 2000.0
            t_cta = max(mma, load) + epi
            waves = _cdiv(n_cta, NUM_SMS)
            est = waves * t_cta
            cand = (est, -TY * TZ)
            if best is None or cand < best[0]:
                best = (cand, ConvTile(TX, TY, TZ, NT, n_ntiles, stages, mt, sb, nx * ty * tz, est))
    if best is None:
        raise ValueError(f"no valid conv tiling for X={X} Y={Y} Z={Z} n_out={n_out} ksize={ksize}")
    return best[1]
