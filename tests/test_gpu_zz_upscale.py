"""Mask assembly at resize ratios far from the BASELINE 4x (postprocess to the ORIGINAL image size, eval.py:266): 8x and
16x up-scaling (a 138x138 prototype map to 1100 / 2208 pixels) and down-scaling, all three mask formats, against
the oracle.  Guards the per-detection column range of the kernel's phase B (tests/test_mask_window_bounds.py)."""
import numpy as np
import pytest
import torch

from oracle import yolact_oracle as O
from yolact_b200.output_utils import assemble_masks, unpack_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(1100, 1100), (2208, 1242), (64, 100)])
def test_masks_at_extreme_resize_ratios(size):
    h, w = size
    r = np.random.RandomState(h + w)
    n, ps = 12, 138
    proto = np.maximum(r.standard_normal((ps, ps, 32)), 0).astype(np.float32)
    proto = (proto + np.roll(proto, 1, 0) + np.roll(proto, 1, 1) + np.roll(proto, (2, 3), (0, 1))) / 4
    coef = np.tanh(r.standard_normal((n, 32))).astype(np.float32)
    c = r.uniform(0.1, 0.9, (n, 2))
    wh = r.uniform(0.05, 0.6, (n, 2))
    box = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    box[2] = [-0.05, 0.3, 0.4, 1.1]
    det = {"box": box, "mask": coef, "class": np.zeros(n, np.int64), "score": np.linspace(0.9, 0.1, n).astype(np.float32),
           "proto": proto}
    _, _, boxes_ref, masks_ref = O.postprocess_one(dict(det), w, h, crop_masks=True)
    t = lambda a: torch.from_numpy(a).cuda()
    for fmt in ("f32", "u8", "bits"):
        m, bpx, _ = assemble_masks(t(proto), t(coef), t(box), h, w, True, fmt)
        got = unpack_bits(m, w) if fmt == "bits" else m
        got = got.cpu().numpy().astype(np.float32)
        assert np.array_equal(bpx.cpu().numpy(), boxes_ref)
        assert (got != masks_ref).mean() < 1e-4, (fmt, size)
        # nothing may be set outside the (padded) crop window's pixel footprint
        assert got.sum() > 0


def test_rle_wider_than_one_pass():
    """w > 1024: the RLE kernel walks the columns in several passes of 1024 threads, carrying the run offset and the
    previous column's last pixel across passes."""
    from oracle import eval_oracle as E
    from yolact_b200.eval_utils import mask_run_lengths
    r = np.random.RandomState(5)
    for (h, w) in [(40, 1300), (7, 2500), (33, 1024), (33, 1025)]:
        ms = [(r.rand(h, w) < p).astype(np.uint8) for p in (0.0, 0.01, 0.5, 1.0)]
        blob = np.zeros((h, w), np.uint8)
        blob[h // 4:h // 2 + 1, 1000:1100] = 1          # straddles the pass boundary at column 1024
        ms.append(blob)
        runs = mask_run_lengths(torch.from_numpy(np.stack(ms)).cuda())
        for i, m in enumerate(ms):
            assert runs[i].tolist() == E.rle_counts(m), (h, w, i)
