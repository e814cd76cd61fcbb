"""CPU: the column range the mask kernel evaluates per detection (yolact_b200/csrc/mask.cu, phase B and the 1-bit
path) must contain every output pixel whose bilinear sources touch the crop window, for any resize ratio.
The kernel's fp32 formula is restated here and checked against the oracle's cropped + resized masks
(non-zero anywhere, not only > 0.5), from 0.46x down-scaling to 16x up-scaling."""
import numpy as np
import pytest

from oracle import yolact_oracle as O

f32 = np.float32


def kernel_column_range(box, pw, out_w):
    # sanitize(bx[0], bx[2], pw, padding=1) then the xa / xb expressions of mask.cu
    x1, x2 = f32(box[0]) * f32(pw), f32(box[2]) * f32(pw)
    cx1 = max(f32(min(x1, x2) - f32(1)), f32(0))
    cx2 = min(f32(max(x1, x2) + f32(1)), f32(pw))
    scale = f32(pw) / f32(out_w)
    xa = int(np.floor(f32(f32(cx1 - f32(0.5)) / scale) - f32(0.5))) - 1
    xb = int(np.ceil(f32(f32(np.ceil(cx2) + f32(0.5)) / scale) - f32(0.5))) + 1
    return max(xa, 0), min(xb, out_w)


@pytest.mark.parametrize("ps,size", [(138, 550), (176, 700), (138, 203), (138, 1100), (138, 2200), (138, 64),
                                     (69, 137), (138, 139), (138, 138)])
def test_column_range_is_a_superset_of_the_nonzero_columns(ps, size):
    r = np.random.RandomState(ps * 31 + size)
    n = 60
    proto = np.maximum(r.standard_normal((ps, ps, 8)), 0).astype(f32) + 2.0     # sigmoid ~ 1: masks fill the window
    coef = (np.abs(np.tanh(r.standard_normal((n, 8)))) + 0.5).astype(f32)
    c = r.uniform(-0.1, 1.1, (n, 2))
    wh = r.uniform(0.001, 0.7, (n, 2))
    box = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(f32)
    box[::7, [0, 2]] = box[::7, [2, 0]]                                       # x1 > x2: sanitize swaps
    nz = O.bilinear_resize(O.proto_masks(proto, coef, box, True), size, size) > 0
    seen = 0
    for i in range(n):
        xa, xb = kernel_column_range(box[i], ps, size)
        cols = np.flatnonzero(nz[i].any(0))
        if cols.size:
            seen += 1
            assert xa <= cols[0] and cols[-1] < xb, (box[i], xa, xb, cols[0], cols[-1])
    assert seen > n // 2
