"""Host logic of sliding-window inference (window grid, importance tables, rank sharding) and the oracle's invariants."""
import pytest
import torch

import mmseg_b200  # noqa: F401
from mmseg_b200.src.trainer import inference as INF
from oracle import sliding_window as OS


def test_window_grid_matches_appendix_c():
    s = OS.window_starts((512, 512, 300), (96, 96, 96), 0.5)
    assert len(s) == 600
    ax0 = sorted({a for a, _, _ in s})
    ax2 = sorted({c for _, _, c in s})
    assert ax0 == [0, 48, 96, 144, 192, 240, 288, 336, 384, 416]
    assert ax2 == [0, 48, 96, 144, 192, 204]
    assert s[0] == (0, 0, 0) and s[1] == (0, 0, 48) and s[6] == (0, 48, 0)  # last axis fastest
    assert len(OS.window_starts((192, 192, 144), (96, 96, 96), 0.5)) == 18
    assert OS.window_starts((128, 128, 128), (96, 96, 96), 0.5)[-1] == (32, 32, 32)
    assert OS.window_starts((96, 96, 96), (96, 96, 96), 0.5) == [(0, 0, 0)]


@pytest.mark.parametrize("size,roi,ov", [((512, 512, 300), (96, 96, 96), 0.5), ((100, 97, 130), (96, 96, 96), 0.25),
                                          ((48, 40, 36), (32, 32, 32), 0.5), ((33, 64, 65), (32, 32, 32), 0.75)])
def test_product_window_grid_equals_oracle(size, roi, ov):
    assert INF.window_starts(size, roi, ov) == OS.window_starts(size, roi, ov)


@pytest.mark.parametrize("mode", ["constant", "gaussian"])
def test_importance_tables_reproduce_oracle_map_bitwise(mode):
    roi = (96, 96, 96)
    tabs, floor = INF.importance_tables(roi, mode)
    w = torch.clamp((tabs[0][:, None, None] * tabs[1][None, :, None]) * tabs[2][None, None, :], min=floor)
    assert torch.equal(w, OS.importance_map(roi, mode))
    if mode == "gaussian":
        assert floor == pytest.approx(1e-3)
        assert (w == floor).float().mean().item() == pytest.approx(0.58, abs=0.02)


def test_oracle_invariants():
    torch.manual_seed(0)
    x = torch.randn(1, 2, 40, 37, 50)
    const = lambda w: torch.full((w.shape[0], 3, *w.shape[2:]), 2.5)
    for mode in ("constant", "gaussian"):
        out, cnt = OS.sliding_window_inference(x, (32, 32, 32), 4, const, 0.5, mode, return_count=True)
        assert torch.allclose(out, torch.full_like(out, 2.5), atol=1e-5)
        assert (cnt > 0).all()
        ident = lambda w: w * 1.0
        out = OS.sliding_window_inference(x, (32, 32, 32), 3, ident, 0.5, mode)
        assert torch.allclose(out, x, atol=1e-5)
    # independent of sw_batch_size
    pred = lambda w: torch.tanh(w) * 2
    a = OS.sliding_window_inference(x, (32, 32, 32), 1, pred, 0.5, "gaussian")
    b = OS.sliding_window_inference(x, (32, 32, 32), 5, pred, 0.5, "gaussian")
    assert torch.equal(a, b)
    # volume smaller than the roi: padded, then cropped back
    small = torch.randn(1, 1, 20, 32, 25)
    out = OS.sliding_window_inference(small, (32, 32, 32), 2, ident, 0.5, "constant")
    assert out.shape == small.shape and torch.allclose(out, small, atol=1e-6)


def test_shard_windows_and_owned_slabs_partition_the_volume():
    starts = OS.window_starts((512, 512, 300), (96, 96, 96), 0.5)
    for world in (1, 2, 3, 4, 8):
        chunks = [INF.shard_windows(len(starts), world, r) for r in range(world)]
        assert chunks[0][0] == 0 and chunks[-1][1] == len(starts)
        assert all(chunks[i][1] == chunks[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in chunks) - min(h - l for l, h in chunks) <= 1
        slabs = INF.owned_slabs(starts, 96, 512, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == 512
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))
    assert INF.touched_range(starts, 0, 75, 96) == (0, 48 + 96)
