"""CPU emulation of conv_tc.cu's data path (TMA box -> flat shifted smem views -> 128-row UMMA tiles -> epilogue).

Used by the `not gpu` tests to check the HOST logic (weight packing, K-chunk tables, tile planner, blocked layout)
against torch.nn.functional on small shapes: it consumes exactly the tensors the kernel would be handed.
It is test infrastructure, not a product path.
"""
import torch


def emulate_conv(src_t, src_cbt, pw, a_cb, tile, n_img, Z, Y, X, garbage=1e4):
    """src_t: float tensor [n_img, cbt, Z, Y, X, 8].  Returns fp32 GEMM output [n_img, n_out, Z, Y, X].

    Mirrors the kernel: accumulator (m, zo) at TMEM column (m*TZ + zo)*NT; for ksize 3 one MMA per (plane, dy, dx, m)
    covers the dz range [dz_lo, dz_hi] = output planes pl-dz_hi .. pl-dz_lo with weight rows stored dz-descending.
    """
    h = pw.ksize // 2
    kt = pw.ksize
    TX, TY, TZ, NT, mt = tile.TX, tile.TY, tile.TZ, tile.NT, tile.mt
    PX, PY = TX + 2 * h, TY + 2 * h
    rows_needed = mt * 128 + 2 * h * PX + 2 * h
    w = pw.w.float()  # [n_ntiles, n_kc, kt*kt, 2, kt*NT, 8]
    out = torch.zeros((n_img, pw.n_out, Z, Y, X))
    for img in range(n_img):
        for z0 in range(0, Z, TZ):
            tz_valid = min(TZ, Z - z0)
            for y0 in range(0, Y, TY):
                for x0 in range(0, X, TX):
                    for nt in range(pw.n_out // NT):
                        tmem = torch.zeros((mt * 128, mt * TZ * NT))  # [lane row within m tile..., columns]
                        acc = torch.zeros((128, mt * TZ * NT))          # lanes x columns, as in TMEM
                        for kc in range(pw.n_kchunks):
                            for pl in range(TZ + 2 * h):
                                z = z0 - h + pl
                                if z < 0 or z >= Z:
                                    continue
                                stage = torch.zeros((2, PY, PX, 8))
                                for kb in range(2):
                                    blk = a_cb[kc] + kb
                                    for yy in range(PY):
                                        y = y0 - h + yy
                                        if y < 0 or y >= Y:
                                            continue
                                        xs0, xs1 = max(x0 - h, 0), min(x0 - h + PX, X)
                                        stage[kb, yy, xs0 - (x0 - h):xs1 - (x0 - h)] = src_t[img, blk, z, y, xs0:xs1]
                                flat = torch.full((2, max(rows_needed, PY * PX), 8), garbage)
                                flat[:, :PY * PX] = stage.reshape(2, PY * PX, 8)
                                dz_hi, dz_lo = min(kt - 1, pl), max(0, pl - tz_valid + 1)
                                if dz_hi < dz_lo:
                                    continue
                                nz = dz_hi - dz_lo + 1
                                r0 = (kt - 1 - dz_hi) * NT
                                c0 = (pl - dz_hi) * NT
                                for dy in range(kt):
                                    for dx in range(kt):
                                        B = w[nt, kc, dy * kt + dx][:, r0:r0 + nz * NT].permute(1, 0, 2).reshape(nz * NT, 16)
                                        for m in range(mt):
                                            a0 = m * 128 + dy * PX + dx
                                            A = flat[:, a0:a0 + 128].permute(1, 0, 2).reshape(128, 16)
                                            cc = m * TZ * NT + c0
                                            acc[:, cc:cc + nz * NT] += A @ B.T
                        for zo in range(tz_valid):
                            z = z0 + zo
                            for m in range(mt):
                                cc = (m * TZ + zo) * NT
                                for r in range(128):
                                    L = m * 128 + r
                                    yy, xx = divmod(L, PX)
                                    y, x = y0 + yy, x0 + xx
                                    if xx < TX and yy < TY and x < X and y < Y:
                                        out[img, nt * NT:(nt + 1) * NT, z, y, x] = acc[r, cc:cc + NT]
    return out


def to_blocked(x, cbt=None):
    """NCDHW float -> [n, cb, Z, Y, X, 8] (channels zero-padded to a multiple of 16)."""
    n, c, Z, Y, X = x.shape
    cp = (c + 15) // 16 * 16
    xp = torch.zeros((n, cp, Z, Y, X), dtype=x.dtype)
    xp[:, :c] = x
    return xp.reshape(n, cp // 8, 8, Z, Y, X).permute(0, 1, 3, 4, 5, 2).contiguous()


# ------------------------------------------------------------------------------------------------ rolling-z schedule
def roll_schedule(Z, ZS, n_kchunks, kpb=1, R=16, n_items_xy=1):
    """The stage stream conv3d_roll_kernel's MMA lane generates for one (x, y) column of a volume of depth Z cut into
    z segments of length ZS (conv_tc.cu, `gen()`): a list of dicts, one per TMA stage, with
      item, q (input plane index inside the segment, z = zs0 - 1 + q), kc (first K chunk of the stage), chunks,
      mma = [(ring slot of the first output plane, number of output planes, first dz), ...]  (1 entry, 2 at a ring wrap),
      acquire = global output-plane indices whose ring slot must be free before the stage (z_empty waits),
      complete = global output-plane indices signalled complete after the stage (z_full commits).
    Pure integer logic restated from the kernel so that the `not gpu` tests can check its invariants."""
    assert n_kchunks % kpb == 0
    n_st = n_kchunks // kpb
    stages, gz, g_acq = [], 0, 0
    n_seg = -(-Z // ZS)
    for item in range(n_seg * n_items_xy):
        zs0 = (item % n_seg) * ZS
        zsv = min(ZS, Z - zs0)
        q_first = 1 if zs0 == 0 else 0
        q_last = zsv if zs0 + zsv >= Z else zsv + 1
        for q in range(q_first, q_last + 1):
            dz_hi = min(2, q)
            dz_lo = max(0, q - (zsv - 1))
            n = dz_hi - dz_lo + 1
            g_lo = gz + (q - dz_hi)
            col = g_lo % R
            n1 = min(n, R - col)
            mma = [(col, n1, dz_hi)] + ([(0, n - n1, dz_hi - n1)] if n > n1 else [])
            zneed = gz + min(q, zsv - 1) + 1
            for kc in range(n_st):
                acquire = list(range(g_acq, zneed)) if g_acq < zneed else []
                g_acq = max(g_acq, zneed)
                complete = []
                if kc == n_st - 1:
                    if q >= 2:
                        complete.append(gz + q - 2)
                    if q == q_last and q_last == zsv:
                        complete.append(gz + zsv - 1)
                stages.append(dict(item=item, q=q, z=zs0 - 1 + q, kc=kc * kpb, chunks=kpb, mma=mma, acquire=acquire,
                                   complete=complete, gz=gz, zsv=zsv, zs0=zs0))
        gz += zsv
    return stages
