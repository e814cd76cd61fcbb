"""Detect (Fast NMS / cross-class) and postprocess (mask assembly) through the C ABI, against the
reference's golden outputs and against the oracle on seeded inputs, plus size-independent properties
at BASELINE sizes (P = 19248 / 57744, 100 detections at 550x550)."""
import numpy as np
import pytest
import torch

from oracle import yolact_oracle as O
from tests.conftest import load_golden
from tests.helpers import unpack_masks
from yolact_b200.detection import Detect
from yolact_b200.output_utils import assemble_masks, postprocess, unpack_bits

pytestmark = pytest.mark.gpu


def _detect(loc, conf, mask, priors, cc=False, logits=False):
    d = Detect(81, bkg_label=0, top_k=200, conf_thresh=0.05, nms_thresh=0.5)
    d.use_fast_nms, d.use_cross_class_nms = True, cc
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    box, coef, cls, score, count = d.detect_padded(t(loc), t(conf), t(mask), t(priors), conf_is_logits=logits)
    torch.cuda.synchronize()
    return [x.cpu().numpy() for x in (box, coef, cls, score, count)]


def test_detect_unit_golden():
    g = load_golden("detect_unit")
    for tag, cc in (("fast", False), ("cc", True)):
        box, coef, cls, score, count = _detect(g["loc"], g["conf"], g["mask"], g["priors"], cc)
        for b in range(2):
            n = int(count[b])
            assert n == g["%s%d_score" % (tag, b)].shape[0]
            assert np.array_equal(cls[b, :n], g["%s%d_class" % (tag, b)])          # bit-exact class ids
            assert np.array_equal(score[b, :n], g["%s%d_score" % (tag, b)])        # scores are copied, exact
            np.testing.assert_allclose(box[b, :n], g["%s%d_box" % (tag, b)], rtol=0, atol=2e-6)
            assert np.array_equal(coef[b, :n], g["%s%d_mask" % (tag, b)])


def _sparse_fixture(seed=5, P=3000, C=81):
    """Few confident priors: fast_nms then returns sub-threshold rows unless second_threshold is on."""
    r = np.random.RandomState(seed)
    priors = np.concatenate([r.uniform(0.1, 0.9, (P, 2)), r.uniform(0.05, 0.3, (P, 2))], 1).astype(np.float32)
    loc = (r.standard_normal((P, 4)) * 0.5).astype(np.float32)
    logits = r.standard_normal((P, C)).astype(np.float32)
    logits[:, 0] += 4.0
    hot = r.choice(P, 40, replace=False)
    logits[hot, r.randint(1, C, 40)] += 7.0
    return loc, O.softmax_rows(logits), r.standard_normal((P, 32)).astype(np.float32), priors


@pytest.mark.parametrize("thresh", [0.05, 0.2])
def test_fast_nms_second_threshold_matches_oracle(thresh):
    """fast_nms(second_threshold=True) (detection.py:137,160-161): a kept row must also have its own class score above
    conf_thresh.  Candidates are selected on the max over classes, so the class rows hold many sub-threshold scores;
    at thresh = 0.2 the plain result carries 42 of them and the flag removes exactly those."""
    loc, conf, mask, priors = _sparse_fixture()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    res = {}
    for flag in (False, True):
        d = Detect(81, bkg_label=0, top_k=200, conf_thresh=thresh, nms_thresh=0.5)
        d.use_fast_nms, d.second_threshold = True, flag
        box, coef, cls, score, count = [x.cpu().numpy() for x in
                                        d.detect_padded(t(loc[None]), t(conf[None]), t(mask[None]), t(priors))]
        want = O.detect_one(loc, conf, mask, priors, conf_thresh=thresh, second_threshold=flag)
        n = int(count[0])
        assert n == want["score"].shape[0]
        assert np.array_equal(cls[0, :n], want["class"]) and np.array_equal(score[0, :n], want["score"])
        np.testing.assert_allclose(box[0, :n], want["box"], rtol=0, atol=2e-6)
        res[flag] = score[0, :n]
    assert (res[True] > thresh).all()
    if thresh == 0.2:
        assert (res[False] <= thresh).sum() > 0 and len(res[True]) < len(res[False])


def test_detect_api_object():
    g = load_golden("detect_unit")
    d = Detect(81, 0, 200, 0.05, 0.5)
    d.use_fast_nms = True   # eval.py:871
    t = lambda a: torch.from_numpy(a).cuda()
    out = d({"loc": t(g["loc"]), "conf": t(g["conf"]), "mask": t(g["mask"]), "priors": t(g["priors"])}, "NET")
    assert len(out) == 2 and out[0]["net"] == "NET"
    det = out[1]["detection"]
    assert det["class"].dtype == torch.int64 and det["box"].shape == (100, 4)
    assert np.array_equal(det["class"].cpu().numpy(), g["fast1_class"])
    with pytest.raises(ValueError):
        Detect(81, 0, 200, 0.05, 0.0)


@pytest.mark.parametrize("case", ["net_resnet50_160", "net_base_192x160_b2", "net_darknet53_160"])
def test_detect_fused_softmax_on_golden_logits(case):
    g = load_golden(case)
    box, coef, cls, score, count = _detect(g["raw_loc"], g["raw_conf"], g["raw_mask"], g["raw_priors"], logits=True)
    for b in range(g["x"].shape[0]):
        n = int(count[b])
        assert n == int(g["det_counts"][b])
        # softmax on the GPU differs from the CPU's in the last ulp -> compare as score-sorted sets
        assert np.array_equal(cls[b, :n], g["det%d_class" % b])
        np.testing.assert_allclose(score[b, :n], g["det%d_score" % b], rtol=0, atol=3e-6)
        np.testing.assert_allclose(box[b, :n], g["det%d_box" % b], rtol=0, atol=2e-6)


def test_detect_none_when_nothing_passes():
    P = 1000
    conf = np.full((1, P, 81), 1.0 / 81, np.float32)
    box, coef, cls, score, count = _detect(np.zeros((1, P, 4), np.float32), conf, np.zeros((1, P, 32), np.float32),
                                           np.tile(np.array([[.5, .5, .1, .1]], np.float32), (P, 1)))
    assert int(count[0]) == 0


def test_detect_ties_are_ordered_by_prior_index():
    # many identical scores: the contract is lower prior index first
    P = 600
    r = np.random.RandomState(0)
    conf = np.zeros((1, P, 81), np.float32)
    conf[:, :, 0] = 0.4
    conf[:, :, 5] = 0.6
    pri = np.concatenate([r.uniform(0.05, 0.95, (P, 2)), np.full((P, 2), 0.01)], 1).astype(np.float32)
    loc = np.zeros((1, P, 4), np.float32)
    mask = r.standard_normal((1, P, 32)).astype(np.float32)
    box, coef, cls, score, count = _detect(loc, conf, mask, pri)
    ref = O.detect_one(loc[0], conf[0], mask[0], pri)
    n = int(count[0])
    assert n == ref["score"].shape[0] == 100
    assert np.array_equal(cls[0, :n], ref["class"])
    np.testing.assert_allclose(box[0, :n], ref["box"], atol=2e-6)
    assert np.array_equal(coef[0, :n], ref["mask"])


@pytest.mark.parametrize("P,seed", [(19248, 0), (57744, 1)])
def test_detect_full_size_vs_oracle(P, seed):
    r = np.random.RandomState(seed)
    logits = (r.standard_normal((1, P, 81)) * 2).astype(np.float32)
    logits[:, :, 0] += 4.0
    pri = np.concatenate([r.uniform(0.05, 0.95, (P, 2)), r.uniform(0.02, 0.5, (P, 2))], 1).astype(np.float32)
    loc = r.standard_normal((1, P, 4)).astype(np.float32)
    mask = np.tanh(r.standard_normal((1, P, 32))).astype(np.float32)
    conf = O.softmax_rows(logits)
    box, coef, cls, score, count = _detect(loc, conf, mask, pri)
    ref = O.detect_one(loc[0], conf[0], mask[0], pri)
    n = int(count[0])
    assert n == ref["score"].shape[0]
    assert np.all(np.diff(score[0, :n]) <= 0)                     # sorted
    assert np.array_equal(cls[0, :n], ref["class"])
    assert np.array_equal(score[0, :n], ref["score"])
    np.testing.assert_allclose(box[0, :n], ref["box"], atol=2e-6)


def test_postprocess_unit_golden_all_formats():
    g = load_golden("postprocess_unit")
    t = lambda a: torch.from_numpy(a).cuda()
    for (h, w) in ((550, 550), (203, 277), (64, 96)):
        for crop in (True, False):
            tag = "%dx%d_%s" % (h, w, "crop" if crop else "nocrop")
            ref = unpack_masks(g["masks_" + tag], w)
            m32, boxes, _ = assemble_masks(t(g["proto"]), t(g["coef"]), t(g["box"]), h, w, crop, "f32")
            mu8, _, _ = assemble_masks(t(g["proto"]), t(g["coef"]), t(g["box"]), h, w, crop, "u8")
            mb, _, _ = assemble_masks(t(g["proto"]), t(g["coef"]), t(g["box"]), h, w, crop, "bits")
            m32 = m32.cpu().numpy()
            assert set(np.unique(m32)) <= {0.0, 1.0}
            assert (m32 != ref).mean() < 1e-4, tag                       # binarised; ulp-level flips only
            assert np.array_equal(boxes.cpu().numpy(), g["boxes_" + tag])  # int64 boxes: exact
            assert np.array_equal(mu8.cpu().numpy().astype(np.float32), m32)
            assert np.array_equal(unpack_bits(mb, w).cpu().numpy().astype(np.float32), m32)


def test_postprocess_api_and_empty():
    g = load_golden("postprocess_unit")
    t = lambda a: torch.from_numpy(a).cuda()
    det = {"box": t(g["box"]), "mask": t(g["coef"]), "class": t(g["cls"]), "score": t(g["score"]), "proto": t(g["proto"])}
    box_before = det["box"].clone()
    classes, scores, boxes, masks = postprocess([{"detection": det, "net": None}], 277, 203)
    assert masks.shape == (23, 203, 277) and masks.dtype == torch.float32 and boxes.dtype == torch.int64
    assert torch.equal(det["box"], box_before)                            # input not mutated
    assert np.array_equal(boxes.cpu().numpy(), g["boxes_203x277_crop"])
    out = postprocess([{"detection": None, "net": None}], 100, 100)
    assert len(out) == 4 and all(o.numel() == 0 for o in out)
    classes, scores, boxes, masks = postprocess([{"detection": dict(det), "net": None}], 96, 64, score_threshold=0.5)
    assert scores.numel() == int((g["score"] > 0.5).sum()) and masks.shape[0] == scores.numel()


def test_postprocess_full_size_properties():
    # BASELINE size: 100 detections at 550x550 from a 138x138x32 prototype tensor
    r = np.random.RandomState(2)
    n = 100
    proto = np.maximum(r.standard_normal((138, 138, 32)), 0).astype(np.float32)
    coef = np.tanh(r.standard_normal((n, 32))).astype(np.float32)
    c, wh = r.uniform(0.2, 0.8, (n, 2)), r.uniform(0.05, 0.5, (n, 2))
    box = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    t = lambda a: torch.from_numpy(a).cuda()
    m, boxes, pm = assemble_masks(t(proto), t(coef), t(box), 550, 550, True, "f32", want_proto_masks=True)
    m = m.cpu().numpy()
    assert m.shape == (n, 550, 550) and set(np.unique(m)) <= {0.0, 1.0}
    # crop property: nothing outside the (padded, upsampled) box; 4 px = one prototype cell of slack + bilinear reach
    bx = boxes.cpu().numpy()
    for i in range(0, n, 7):
        x1, y1, x2, y2 = bx[i]
        outside = m[i].copy()
        outside[max(0, y1 - 10):y2 + 10, max(0, x1 - 10):x2 + 10] = 0
        assert outside.sum() == 0
    # against the oracle on a subset
    det = {"box": box[:8], "mask": coef[:8], "class": np.zeros(8, np.int64), "score": np.ones(8, np.float32), "proto": proto}
    _, _, ob, om = O.postprocess_one(det, 550, 550)
    assert np.array_equal(ob, bx[:8])
    assert (om != m[:8]).mean() < 1e-4
    np.testing.assert_allclose(pm.cpu().numpy()[:8], O.proto_masks(proto, coef[:8], box[:8]), atol=2e-6)


def test_postprocess_batch_equals_per_image():
    from yolact_b200.output_utils import assemble_masks_batch
    r = np.random.RandomState(11)
    B, n = 3, 17
    proto = np.maximum(r.standard_normal((B, 40, 44, 32)), 0).astype(np.float32)
    coef = np.tanh(r.standard_normal((B, n, 32))).astype(np.float32)
    c, wh = r.uniform(0.2, 0.8, (B, n, 2)), r.uniform(0.05, 0.6, (B, n, 2))
    box = np.concatenate([c - wh / 2, c + wh / 2], 2).astype(np.float32)
    t = lambda a: torch.from_numpy(a).cuda()
    for fmt in ("f32", "u8", "bits"):
        mb, bb = assemble_masks_batch(t(proto), t(coef), t(box), 101, 135, True, fmt)
        for b in range(B):
            m1, b1, _ = assemble_masks(t(proto[b]), t(coef[b]), t(box[b]), 101, 135, True, fmt)
            assert torch.equal(mb[b], m1) and torch.equal(bb[b], b1), (fmt, b)
