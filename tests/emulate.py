"""CPU emulation of conv_tc.cu's data path (TMA box -> flat shifted smem views -> 128-row UMMA tiles -> epilogue).

Used by the `not gpu` tests to check the HOST logic (weight packing, K-chunk tables, tile planner, blocked layout)
against torch.nn.functional on small shapes: it consumes exactly the tensors the kernel would be handed.
It is test infrastructure, not a product path.
"""
import torch


def emulate_conv(src_t, src_cbt, pw, a_cb, tile, n_img, Z, Y, X, garbage=1e4):
    """src_t: float tensor [n_img, cbt, Z, Y, X, 8].  Returns fp32 GEMM output [n_img, n_out, Z, Y, X]."""
    h = pw.ksize // 2
    kt = pw.ksize
    TX, TY, TZ, NT, mt = tile.TX, tile.TY, tile.TZ, tile.NT, tile.mt
    PX, PY = TX + 2 * h, TY + 2 * h
    rows_needed = mt * 128 + 2 * h * PX + 2 * h
    w = pw.w.float()  # [n_ntiles, n_kc, taps, 2, NT, 8]
    out = torch.zeros((n_img, pw.n_out, Z, Y, X))
    g = torch.Generator().manual_seed(1)
    for img in range(n_img):
        for z0 in range(0, Z, TZ):
            for y0 in range(0, Y, TY):
                for x0 in range(0, X, TX):
                    for nt in range(pw.n_out // NT):
                        acc = torch.zeros((TZ, mt * 128, NT))
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
                                for dz in range(kt):
                                    zo = pl - dz
                                    if zo < 0 or zo >= TZ or z0 + zo >= Z:
                                        continue
                                    for dy in range(kt):
                                        for dx in range(kt):
                                            tap = (dz * kt + dy) * kt + dx
                                            B = w[nt, kc, tap].permute(1, 0, 2).reshape(NT, 16)
                                            for m in range(mt):
                                                r0 = m * 128 + dy * PX + dx
                                                A = flat[:, r0:r0 + 128].permute(1, 0, 2).reshape(128, 16)
                                                acc[zo, m * 128:(m + 1) * 128] += A @ B.T
                        for zo in range(TZ):
                            z = z0 + zo
                            if z >= Z:
                                break
                            for L in range(mt * 128):
                                yy, xx = divmod(L, PX)
                                y, x = y0 + yy, x0 + xx
                                if xx < TX and yy < TY and x < X and y < Y:
                                    out[img, nt * NT:(nt + 1) * NT, z, y, x] = acc[zo, L]
    return out


def to_blocked(x, cbt=None):
    """NCDHW float -> [n, cb, Z, Y, X, 8] (channels zero-padded to a multiple of 16)."""
    n, c, Z, Y, X = x.shape
    cp = (c + 15) // 16 * 16
    xp = torch.zeros((n, cp, Z, Y, X), dtype=x.dtype)
    xp[:, :c] = x
    return xp.reshape(n, cp // 8, 8, Z, Y, X).permute(0, 1, 3, 4, 5, 2).contiguous()
