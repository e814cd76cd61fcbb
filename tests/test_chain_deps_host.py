"""Chain kernel, host side (no GPU): the dependency arithmetic the kernel uses to decide which tiles of the producing
layer a tile has to wait for (csrc/tc_conv.cu: chain_rows_of_tile / chain_input_rows / chain_tiles_of_rows, exposed
through yb_debug_chain_deps) against a brute-force restatement -- every input pixel a consumer tile reads must lie in a
producer tile inside the returned range, for flattened and 2-D tilings on both sides, strides, padding, odd sizes and
batches (tiles that straddle rows and images).
"""
import ctypes
import itertools

import numpy as np
import pytest

from yolact_b200 import _lib


def deps(B, Hin, Win, k, stride, pad, producer_flat, m):
    out = (ctypes.c_int32 * 14)()
    _lib.check(_lib.load().yb_debug_chain_deps(B, Hin, Win, k, stride, pad, int(producer_flat), m, out), "yb_debug_chain_deps")
    return list(out)


def tile_pixels(flat, tw, th, tiles_x, tiles_y, B, H, W, m):
    """(b, y, x) of the valid output pixels of M tile m."""
    if flat:
        lo, hi = m * tw, min(m * tw + tw, B * H * W)
        idx = np.arange(lo, hi)
        return idx // (H * W), (idx % (H * W)) // W, idx % W
    tx, ty, b = m % tiles_x, (m // tiles_x) % tiles_y, m // (tiles_x * tiles_y)
    ys = np.arange(ty * th, min(ty * th + th, H))
    xs = np.arange(tx * tw, min(tx * tw + tw, W))
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    return np.full(yy.size, b), yy.ravel(), xx.ravel()


def tile_of_pixel(flat, tw, th, tiles_x, tiles_y, H, W, b, y, x):
    if flat:
        return (b * H * W + y * W + x) // tw
    return (b * tiles_y + y // th) * tiles_x + x // tw


GEOMS = [  # B, Hin, Win
    (8, 35, 35), (8, 69, 69), (2, 138, 138), (3, 18, 18), (1, 5, 5), (3, 32, 40), (5, 9, 13), (2, 44, 44), (1, 1, 7), (4, 23, 23),
]
LAYERS = [(1, 1, 0), (3, 1, 1), (3, 2, 1), (1, 2, 0)]   # k, stride, pad: the bottleneck's four kinds of layers


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("layer", LAYERS)
@pytest.mark.parametrize("producer_flat", [0, 1])
def test_every_input_pixel_is_covered(geom, layer, producer_flat):
    B, Hin, Win = geom
    k, s, p = layer
    Ho, Wo = (Hin + 2 * p - k) // s + 1, (Win + 2 * p - k) // s + 1
    d0 = deps(B, Hin, Win, k, s, p, producer_flat, 0)
    cflat, ctw, cth, ctx, cty, cm = d0[0:6]
    pflat, ptw, pth, ptx, pty, pm = d0[6:12]
    assert cflat == int(k == 1 and s == 1 and p == 0) and pflat == producer_flat
    assert ctw * cth <= 128 and ptw * pth <= 128
    assert cm == (ctx * cty if cflat else ctx * cty * B) and pm == (ptx * pty if pflat else ptx * pty * B)
    covered = np.zeros(cm, bool)
    for m in range(cm):
        d = deps(B, Hin, Win, k, s, p, producer_flat, m)
        first, last = d[12], d[13]
        assert 0 <= first <= last < pm, (m, first, last, pm)
        b, y, x = tile_pixels(cflat, ctw, cth, ctx, cty, B, Ho, Wo, m)
        assert b.size > 0
        covered[m] = True
        need = set()
        for dy, dx in itertools.product(range(k), range(k)):
            iy, ix = y * s - p + dy, x * s - p + dx
            ok = (iy >= 0) & (iy < Hin) & (ix >= 0) & (ix < Win)
            t = tile_of_pixel(pflat, ptw, pth, ptx, pty, Hin, Win, b[ok], iy[ok], ix[ok])
            need.update(np.unique(t).tolist())
        assert need, m
        assert min(need) >= first and max(need) <= last, (m, sorted(need)[:4], sorted(need)[-4:], first, last)
        # and the range is not grossly conservative: at most the tiles of two extra rows' worth on each side
        slack = 2 * (Win // max(ptw, 1) + 2) * (1 if pflat else ptx)
        assert first >= min(need) - slack and last <= max(need) + slack, (m, min(need), max(need), first, last)
    assert covered.all()


def test_bad_arguments_are_rejected():
    out = (ctypes.c_int32 * 14)()
    lib = _lib.load()
    assert lib.yb_debug_chain_deps(1, 8, 8, 5, 1, 2, 0, 0, out) != 0          # 5x5: not a chain layer
    assert lib.yb_debug_chain_deps(1, 8, 8, 3, 1, 1, 0, 10 ** 6, out) != 0    # tile out of range


def _check_tile(B, Hin, Win, k, s, p, producer_flat, m, d0):
    cflat, ctw, cth, ctx, cty, cm = d0[0:6]
    pflat, ptw, pth, ptx, pty, pm = d0[6:12]
    Ho, Wo = (Hin + 2 * p - k) // s + 1, (Win + 2 * p - k) // s + 1
    d = deps(B, Hin, Win, k, s, p, producer_flat, m)
    first, last = d[12], d[13]
    assert 0 <= first <= last < pm
    b, y, x = tile_pixels(cflat, ctw, cth, ctx, cty, B, Ho, Wo, m)
    need = set()
    for dy, dx in itertools.product(range(k), range(k)):
        iy, ix = y * s - p + dy, x * s - p + dx
        ok = (iy >= 0) & (iy < Hin) & (ix >= 0) & (ix < Win)
        need.update(np.unique(tile_of_pixel(pflat, ptw, pth, ptx, pty, Hin, Win, b[ok], iy[ok], ix[ok])).tolist())
    assert need and min(need) >= first and max(need) <= last, (B, Hin, Win, k, s, p, producer_flat, m, first, last)


def test_random_geometries():
    """Hypothesis sweep over batch / image sizes the fixed grid above does not contain (first, middle and last tiles)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(B=st.integers(1, 9), Hin=st.integers(1, 150), Win=st.integers(1, 150), layer=st.sampled_from(LAYERS),
           producer_flat=st.integers(0, 1), pick=st.floats(0, 1))
    def run(B, Hin, Win, layer, producer_flat, pick):
        k, s, p = layer
        if (Hin + 2 * p - k) // s + 1 < 1 or (Win + 2 * p - k) // s + 1 < 1:
            return
        d0 = deps(B, Hin, Win, k, s, p, producer_flat, 0)
        cm = d0[5]
        for m in sorted({0, cm - 1, int(pick * (cm - 1))}):
            _check_tile(B, Hin, Win, k, s, p, producer_flat, m, d0)

    run()
