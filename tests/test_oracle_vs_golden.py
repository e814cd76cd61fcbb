"""Pins oracle/yolact_oracle.py (the CPU restatement) to the REAL reference through the committed
golden vectors (tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import yolact_oracle as O
from oracle.weights import deterministic_state_dict
from tests.helpers import cfg_for, unpack_masks, rel_err
from tests.conftest import load_golden

NET_CASES = ["net_resnet50_160", "net_base_192x160_b2", "net_plus_resnet50_256", "net_darknet53_160"]


def _oracle_net(case):
    import yolact_b200
    g = load_golden(case)
    cfg = cfg_for(str(g["config"]))
    net = yolact_b200.Yolact(cfg)
    sd = deterministic_state_dict(net.state_dict(), int(g["seed"]))
    return g, cfg, O.ConvStackOracle(cfg, sd)


@pytest.mark.parametrize("case", NET_CASES)
def test_conv_stack_matches_reference(case):
    g, cfg, orc = _oracle_net(case)
    out = orc.forward(torch.from_numpy(g["x"]))
    rs = int(g["row_stride"])
    # fp32 CPU vs fp32 CPU: only summation-order noise
    assert np.array_equal(out["priors"].numpy(), g["raw_priors"])
    assert rel_err(out["proto"].numpy(), g["raw_proto"]) < 2e-5
    assert rel_err(out["loc"].numpy()[:, ::rs], g["raw_loc"]) < 2e-5
    assert rel_err(out["conf"].numpy()[:, ::rs], g["raw_conf"]) < 2e-5
    assert rel_err(out["mask"].numpy()[:, ::rs], g["raw_mask"]) < 2e-5


@pytest.mark.parametrize("case", ["net_resnet50_160", "net_base_192x160_b2", "net_darknet53_160"])
def test_detect_and_postprocess_match_reference(case):
    g, cfg, orc = _oracle_net(case)
    conf = O.softmax_rows(g["raw_conf"])
    for b in range(g["x"].shape[0]):
        det = O.detect_one(g["raw_loc"][b], conf[b], g["raw_mask"][b], g["raw_priors"])
        assert det is not None and det["score"].shape[0] == int(g["det_counts"][b])
        assert np.array_equal(det["class"], g["det%d_class" % b])            # class ids: exact
        np.testing.assert_allclose(det["score"], g["det%d_score" % b], rtol=0, atol=2e-6)
        np.testing.assert_allclose(det["box"], g["det%d_box" % b], rtol=0, atol=2e-6)
        np.testing.assert_allclose(det["mask"], g["det%d_mask" % b], rtol=0, atol=1e-6)
        ph, pw = (int(v) for v in g["post_hw"])
        det["proto"] = g["raw_proto"][b]
        classes, scores, boxes, masks = O.postprocess_one(det, pw, ph)
        ref = unpack_masks(g["post%d_masks_packed" % b], pw)
        assert np.array_equal(boxes, g["post%d_boxes" % b])
        assert (masks != ref).mean() < 1e-4                                   # binarised: a few ulp-level flips at most


@pytest.mark.parametrize("tag", ["full_base_550_to_480x640", "full_plus_resnet50_550", "full_im700_700"])
def test_oracle_matches_reference_at_full_size(tag):
    """BASELINE.json configs at FULL resolution (550 / 700 px, DCN at C = 128..512, 640x480 postprocess target): the
    oracle's whole pipeline against the real reference's outputs.  Ranks may swap only between scores closer than
    2e-5 (fp32 summation order decides those in any implementation; torch's own CPU conv is not bit-reproducible
    across thread counts)."""
    from oracle import torch_port as T
    from oracle.weights import deterministic_input
    from tests.fullsize_golden import load_case, raw_errors
    from tests.parity_utils import align
    import yolact_b200
    g, cfg, ref, (ph, pw) = load_case(tag)
    sd = deterministic_state_dict(yolact_b200.Yolact(cfg).state_dict(), int(g["seed"]))
    orc = O.ConvStackOracle(cfg, sd)
    x = deterministic_input(1, int(g["size"]), int(g["size"]), int(g["seed_x"]))
    raw = orc.forward(x)
    e = raw_errors({k: v.numpy() for k, v in raw.items()}, g)
    assert e["priors_equal"]
    for k in ("loc", "conf", "mask", "proto"):
        assert e[k] < 5e-5, (k, e[k])
    det = T.detect_one(raw["loc"][0], torch.softmax(raw["conf"], -1)[0], raw["mask"][0], raw["priors"],
                       cfg.nms_conf_thresh, cfg.nms_thresh, cfg.nms_top_k, cfg.max_num_detections)
    det["proto"] = raw["proto"][0]
    box_rel = det["box"].clone().numpy()
    classes, scores, boxes, masks = T.postprocess_one(det, pw, ph, maskiou_fn=orc.maskiou if cfg.use_maskiou else None)
    got = {"class": classes.numpy(), "score": (scores[0] if isinstance(scores, list) else scores).numpy(), "box": box_rel}
    perm, ok = align(got, ref)
    assert ok, "class ids differ beyond score ties"
    assert np.abs(got["box"][perm] - ref["box"]).max() < 1e-5 and np.abs(got["score"][perm] - ref["score"]).max() < 1e-5
    assert np.abs(boxes.numpy()[perm] - ref["box_px"]).max() <= 1
    assert ((masks.numpy()[perm] > 0.5) != ref["masks"]).mean() < 1e-4
    if ref["score_maskiou"] is not None:
        np.testing.assert_allclose(scores[1].numpy()[perm], ref["score_maskiou"], rtol=1e-3, atol=1e-5)


def test_detect_unit_fast_and_cross_class():
    g = load_golden("detect_unit")
    for b in range(2):
        for tag, cc in (("fast", False), ("cc", True)):
            det = O.detect_one(g["loc"][b], g["conf"][b], g["mask"][b], g["priors"], cross_class=cc)
            assert np.array_equal(det["class"], g["%s%d_class" % (tag, b)])
            np.testing.assert_allclose(det["score"], g["%s%d_score" % (tag, b)], rtol=0, atol=1e-6)
            np.testing.assert_allclose(det["box"], g["%s%d_box" % (tag, b)], rtol=0, atol=2e-6)
            np.testing.assert_allclose(det["mask"], g["%s%d_mask" % (tag, b)], rtol=0, atol=0)


def test_postprocess_unit():
    g = load_golden("postprocess_unit")
    det = {"box": g["box"], "mask": g["coef"], "class": g["cls"], "score": g["score"], "proto": g["proto"]}
    for (h, w) in ((550, 550), (203, 277), (64, 96)):
        for crop in (True, False):
            tag = "%dx%d_%s" % (h, w, "crop" if crop else "nocrop")
            classes, scores, boxes, masks = O.postprocess_one(dict(det), w, h, crop_masks=crop)
            assert np.array_equal(boxes, g["boxes_" + tag])
            ref = unpack_masks(g["masks_" + tag], w)
            assert (masks != ref).mean() < 1e-4, tag


def test_plus_maskiou_scores():
    g, cfg, orc = _oracle_net("net_plus_resnet50_256")
    b = 0
    det = {k: g["det%d_%s" % (b, k)] for k in ("box", "mask", "class", "score")}
    det["proto"] = g["raw_proto"][b]
    fn = lambda pm: orc.maskiou(torch.from_numpy(pm).unsqueeze(1)).numpy()
    ph, pw = (int(v) for v in g["post_hw"])
    classes, scores, boxes, masks = O.postprocess_one(det, pw, ph, maskiou_fn=fn)
    assert isinstance(scores, list)                                           # output_utils.py:84-88
    np.testing.assert_allclose(scores[0], g["post0_scores"], atol=1e-6)
    np.testing.assert_allclose(scores[1], g["post0_scores_maskiou"], rtol=2e-4, atol=1e-6)
    assert (masks != unpack_masks(g["post0_masks_packed"], pw)).mean() < 1e-4


def test_dcn_restatement():
    g = load_golden("dcn_unit")
    for tag in ("s1", "s2"):
        y = O.dcn_v2_forward(g[tag + "_x"], g[tag + "_offset"], g[tag + "_mask"], g[tag + "_w"], g[tag + "_bias"],
                             int(g[tag + "_stride"]), 1, 1)
        assert np.abs(y - g[tag + "_y"]).max() < 2e-5
    # reference's own known answer: zero offsets, mask 0.5, identity kernel -> 2*out == in (test.py:32-67)
    r = np.random.RandomState(0)
    C = 8
    x = r.standard_normal((2, C, 7, 9)).astype(np.float32)
    w = np.zeros((C, C, 3, 3), np.float32)
    w[np.arange(C), np.arange(C), 1, 1] = 1
    y = O.dcn_v2_forward(x, np.zeros((2, 18, 7, 9), np.float32), np.full((2, 9, 7, 9), 0.5, np.float32), w,
                         np.zeros(C, np.float32), 1, 1, 1)
    assert np.abs(2 * y - x).max() < 1e-10


def test_torch_port_matches_numpy_oracle_and_golden():
    """oracle/torch_port.py (the CPU-baseline implementation bench.py times) against the golden vectors."""
    from oracle import torch_port as T
    g = load_golden("detect_unit")
    t = torch.from_numpy
    for b in range(2):
        det = T.detect_one(t(g["loc"][b]), t(g["conf"][b]), t(g["mask"][b]), t(g["priors"]))
        assert np.array_equal(det["class"].numpy(), g["fast%d_class" % b])
        np.testing.assert_allclose(det["score"].numpy(), g["fast%d_score" % b], atol=1e-6)
        np.testing.assert_allclose(det["box"].numpy(), g["fast%d_box" % b], atol=2e-6)
    g = load_golden("postprocess_unit")
    det = {"box": t(g["box"]), "mask": t(g["coef"]), "class": t(g["cls"]), "score": t(g["score"]), "proto": t(g["proto"])}
    classes, scores, boxes, masks = T.postprocess_one(det, 277, 203)
    assert np.array_equal(boxes.numpy(), g["boxes_203x277_crop"])
    assert (masks.numpy() != unpack_masks(g["masks_203x277_crop"], 277)).mean() < 1e-4
