"""CPU: pins oracle/eval_oracle.py and the traditional-NMS restatement to the REAL reference through
tests/golden/eval_unit.npz / detect_unit.npz (oracle/gen_golden.py), and checks the host-side RLE string
encoder of the package against the oracle's scalar restatement of maskApi.c."""
import numpy as np
import pytest

from oracle import eval_oracle as E
from oracle import yolact_oracle as O
from tests.conftest import load_golden

XF_CASES = ["up", "down", "same", "ar", "dark"]


def xf_mode(cfgrow):
    _, _, normalize, subtract_means, to_float = (int(v) for v in cfgrow)
    return "normalize" if normalize else ("subtract_means" if subtract_means else ("to_float" if to_float else "none"))


@pytest.mark.parametrize("case", XF_CASES)
def test_fast_base_transform_matches_reference(case):
    g = load_golden("eval_unit")
    ref = g["xf_%s_out" % case]
    y = E.fast_base_transform(g["xf_%s_img" % case], ref.shape[2], ref.shape[3], xf_mode(g["xf_%s_cfg" % case]))
    assert y.shape == ref.shape
    # fp32 vs fp32: ATen may contract a*b+c; inputs are 0..255 -> 1e-5 absolute after the /std
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-5)


def test_preserve_aspect_ratio_size():
    from yolact_b200.augmentations import calc_size_preserve_ar
    g = load_golden("eval_unit")
    ref = g["xf_ar_out"]
    img = g["xf_ar_img"]
    w, h = calc_size_preserve_ar(img.shape[2], img.shape[1], int(g["xf_ar_cfg"][0]))
    assert (h, w) == ref.shape[2:]


def test_traditional_nms_matches_reference():
    g = load_golden("detect_unit")
    for ms in (550, 138):
        for b in range(2):
            det = O.detect_one(g["loc"][b], g["conf"][b], g["mask"][b], g["priors"], traditional=True, max_size=ms)
            tag = "trad%d_%d_" % (ms, b)
            assert np.array_equal(det["class"], g[tag + "class"])
            np.testing.assert_allclose(det["score"], g[tag + "score"], rtol=0, atol=1e-6)
            np.testing.assert_allclose(det["box"], g[tag + "box"], rtol=0, atol=2e-6)
            np.testing.assert_allclose(det["mask"], g[tag + "mask"], rtol=0, atol=0)


def test_cython_nms_small_case():
    # hand-checked: box 1 overlaps box 0 by (10*10)/(121+121-100) = 0.704 -> suppressed; box 2 is disjoint
    d = np.array([[0, 0, 10, 10, 0.9], [1, 1, 11, 11, 0.8], [50, 50, 60, 60, 0.7]], np.float32)
    assert O.cython_nms(d, 0.5).tolist() == [0, 2]
    assert O.cython_nms(d, 0.71).tolist() == [0, 1, 2]


def test_mask_and_box_iou_match_reference():
    g = load_golden("eval_unit")
    for crowd, tag in ((False, "plain"), (True, "crowd")):
        a = E.mask_iou(g["iou_masks_a"], g["iou_masks_b"], crowd)
        assert np.array_equal(a, g["iou_mask_" + tag], equal_nan=True)      # integer counts + one division: exact
        b = E.box_iou(g["iou_boxes_a"], g["iou_boxes_b"], crowd)
        np.testing.assert_allclose(b, g["iou_box_" + tag], rtol=0, atol=1e-7)


@pytest.mark.parametrize("tag,kw", [("masks", dict(top_k=8, score_threshold=0.15)),
                                    ("classcolor", dict(top_k=15, score_threshold=0.3, class_color=True))])
def test_prep_display_blend_matches_reference(tag, kw):
    g = load_golden("eval_unit")
    det = {"box": g["disp_box"], "mask": g["disp_coef"], "class": g["disp_cls"], "score": g["disp_score"],
           "proto": g["disp_proto"]}
    out = E.prep_display_masks(det, g["disp_frame"].astype(np.float32), **kw)
    ref = g["disp_" + tag]
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1                      # .byte() truncation of an fp32 sum: +-1 LSB where the order differs
    assert (diff > 0).mean() < 2e-3


def test_rle_roundtrip_and_known_vectors():
    r = np.random.RandomState(0)
    for (h, w) in [(1, 1), (3, 5), (17, 9), (64, 33), (203, 277)]:
        for p in (0.0, 0.02, 0.5, 1.0):
            m = (r.rand(h, w) < p).astype(np.uint8)
            c = E.rle_counts(m)
            assert sum(c) == h * w and (len(c) % 2 == 1) == (not m.T.reshape(-1)[-1])
            assert E.rle_from_string(E.rle_to_string(c)) == c
            assert np.array_equal(E.rle_decode(c, h, w), m)
    # hand-derived: [[0,1],[1,1]] column-major 0,1,1,1 -> counts [1,3] -> chars '1','3'
    assert E.rle_counts(np.array([[0, 1], [1, 1]])) == [1, 3]
    assert E.rle_to_string([1, 3]) == b"13"
    # first pixel set -> leading zero-length run
    assert E.rle_counts(np.array([[1, 0]])) == [0, 1, 1]
    # a value >= 16 needs a continuation char; a negative delta sets the sign bit: 5 -> '5'; 100 -> 'T3'; 3-5=-2 -> 'N'
    assert E.rle_to_string([0, 5, 100, 3]) == b"05T3N"


def test_package_rle_string_matches_oracle():
    from yolact_b200.eval_utils import rle_to_string
    r = np.random.RandomState(1)
    for _ in range(20):
        n = r.randint(1, 400)
        c = r.randint(0, 5000, size=n).tolist()
        assert rle_to_string(np.array(c, np.uint32)) == E.rle_to_string(c)
    big = [0, 1 << 20, 3, (1 << 31) - 7, 1, 2]
    assert rle_to_string(np.array(big, np.uint32)) == E.rle_to_string(big)


def _rle_encode_loop(mask):
    """maskApi.c rleEncode, literally: walk the column-major pixels, emit a count at every value change."""
    flat = np.asarray(mask).T.reshape(-1)
    cnts, c, p = [], 0, 0
    for v in flat:
        v = int(v)
        if v != p:
            cnts.append(c)
            c = 0
            p = v
        c += 1
    cnts.append(c)
    return cnts


def test_vectorised_rle_counts_equal_the_literal_loop():
    r = np.random.RandomState(3)
    for (h, w) in [(1, 1), (1, 7), (6, 1), (5, 4), (13, 17)]:
        for p in (0.0, 0.2, 0.5, 0.9, 1.0):
            m = (r.rand(h, w) < p).astype(np.uint8)
            assert E.rle_counts(m) == _rle_encode_loop(m), (h, w, p)


def test_rle_string_round_trip_property():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=0, max_value=(1 << 31) - 1), min_size=1, max_size=64))
    def check(counts):
        from yolact_b200.eval_utils import rle_to_string
        s = E.rle_to_string(counts)
        assert all(48 <= ch < 48 + 64 for ch in s)               # printable: '0' .. 'o'
        assert E.rle_from_string(s) == counts
        assert rle_to_string(np.array(counts, np.int64)) == s    # the package's vectorised encoder agrees

    check()
