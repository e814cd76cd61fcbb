"""SURVEY.md section 8f rows through the C ABI on the GPU: FastBaseTransform, traditional NMS (--fast_nms=False),
mask_iou / jaccard on bit-packed masks, COCO RLE, prep_display's blend -- against the reference's golden outputs
(tests/golden/eval_unit.npz, detect_unit.npz), against the oracle on seeded inputs, and through size-independent
properties at BASELINE sizes (100 masks at 550x550, P = 19248)."""
import numpy as np
import pytest
import torch

from oracle import eval_oracle as E
from oracle import yolact_oracle as O
from tests.conftest import load_golden
from tests.helpers import cfg_for
from tests.test_eval_rows_oracle import XF_CASES, xf_mode
from yolact_b200 import config as ybcfg
from yolact_b200.augmentations import FastBaseTransform
from yolact_b200.detection import Detect
from yolact_b200.eval_utils import (display_blend, encode_masks, get_color, jaccard, mask_iou, mask_run_lengths,
                                    pack_masks)
from yolact_b200.output_utils import postprocess
from yolact_b200.output_utils import assemble_masks

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- FastBaseTransform ------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", XF_CASES)
@pytest.mark.parametrize("as_u8", [True, False])
def test_fast_base_transform_golden(case, as_u8):
    g = load_golden("eval_unit")
    S, ar, normalize, subtract_means, to_float = (int(v) for v in g["xf_%s_cfg" % case])
    c = cfg_for("yolact_base_config")
    c.max_size, c.preserve_aspect_ratio = S, bool(ar)
    c.normalize, c.subtract_means, c.to_float = bool(normalize), bool(subtract_means), bool(to_float)
    img = cuda(g["xf_%s_img" % case])
    y = FastBaseTransform(c)(img if as_u8 else img.float())
    ref = g["xf_%s_out" % case]
    assert tuple(y.shape) == ref.shape and y.dtype == torch.float32
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=0, atol=2e-5)


def test_fast_base_transform_full_size_vs_oracle_and_properties():
    r = np.random.RandomState(2)
    img = r.randint(0, 256, size=(2, 480, 640, 3)).astype(np.uint8)        # COCO-typical frame -> 550x550
    c = cfg_for("yolact_base_config")
    y = FastBaseTransform(c)(cuda(img)).cpu().numpy()
    ref = E.fast_base_transform(img, 550, 550, "normalize")
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-5)
    # a constant frame stays constant through any resize: (v - mean) / std per channel, RGB order
    flat = np.full((1, 123, 77, 3), 0, np.uint8)
    flat[..., 0], flat[..., 1], flat[..., 2] = 10, 100, 200               # B, G, R
    z = FastBaseTransform(c)(cuda(flat)).cpu().numpy()
    for ch, (v, m, s) in enumerate(zip((200, 100, 10), ybcfg.MEANS[::-1], ybcfg.STD[::-1])):
        np.testing.assert_allclose(z[0, ch], (np.float32(v) - np.float32(m)) / np.float32(s), rtol=0, atol=1e-6)
    with pytest.raises(Exception):
        FastBaseTransform(c)(torch.zeros(1, 8, 8, 3))                     # CPU tensor: no fallback


# ---- traditional NMS ----------------------------------------------------------------------------------------
def _detect_trad(loc, conf, mask, priors, max_size, logits=False):
    c = cfg_for("yolact_base_config")
    c.max_size = max_size
    d = Detect(81, bkg_label=0, top_k=200, conf_thresh=0.05, nms_thresh=0.5, cfg=c)
    d.use_fast_nms = False
    box, coef, cls, score, count = d.detect_padded(cuda(loc), cuda(conf), cuda(mask), cuda(priors), conf_is_logits=logits)
    torch.cuda.synchronize()
    return [x.cpu().numpy() for x in (box, coef, cls, score, count)]


@pytest.mark.parametrize("ms", [550, 138])
def test_traditional_nms_golden(ms):
    g = load_golden("detect_unit")
    box, coef, cls, score, count = _detect_trad(g["loc"], g["conf"], g["mask"], g["priors"], ms)
    for b in range(2):
        tag = "trad%d_%d_" % (ms, b)
        n = int(count[b])
        assert n == g[tag + "score"].shape[0]
        assert np.array_equal(cls[b, :n], g[tag + "class"])                # class ids: exact
        assert np.array_equal(score[b, :n], g[tag + "score"])              # scores are copied: exact
        np.testing.assert_allclose(box[b, :n], g[tag + "box"], rtol=0, atol=2e-6)
        assert np.array_equal(coef[b, :n], g[tag + "mask"])


def test_traditional_nms_many_candidates_vs_oracle():
    # > 256 candidates per class in a few classes: exercises the chunked walk and the kept-box carry
    r = np.random.RandomState(7)
    P = 4000
    pri = np.concatenate([r.uniform(0.1, 0.9, (P, 2)), r.uniform(0.02, 0.15, (P, 2))], 1).astype(np.float32)
    loc = (r.standard_normal((1, P, 4)) * 0.5).astype(np.float32)
    conf = np.zeros((1, P, 81), np.float32)
    hot = r.randint(1, 4, size=P)                                         # only classes 1..3 -> ~1300 candidates each
    sc = r.uniform(0.06, 0.95, size=P).astype(np.float32)
    conf[0, np.arange(P), hot] = sc
    conf[0, :, 0] = 1 - sc
    mask = np.tanh(r.standard_normal((1, P, 32))).astype(np.float32)
    box, coef, cls, score, count = _detect_trad(loc, conf, mask, pri, 550)
    det = O.detect_one(loc[0], conf[0], mask[0], pri, traditional=True, max_size=550)
    n = int(count[0])
    assert n == det["score"].shape[0] == 100
    assert np.array_equal(cls[0, :n], det["class"])
    assert np.array_equal(score[0, :n], det["score"])
    np.testing.assert_allclose(box[0, :n], det["box"], rtol=0, atol=2e-6)
    # property: no kept pair of one class overlaps >= 0.5 under the +1 pixel convention
    b = box[0, :n] * 550
    for c in np.unique(cls[0, :n]):
        bb = b[cls[0, :n] == c]
        for i in range(len(bb)):
            for j in range(i + 1, len(bb)):
                w = max(0.0, min(bb[i, 2], bb[j, 2]) - max(bb[i, 0], bb[j, 0]) + 1)
                h = max(0.0, min(bb[i, 3], bb[j, 3]) - max(bb[i, 1], bb[j, 1]) + 1)
                ai = (bb[i, 2] - bb[i, 0] + 1) * (bb[i, 3] - bb[i, 1] + 1)
                aj = (bb[j, 2] - bb[j, 0] + 1) * (bb[j, 3] - bb[j, 1] + 1)
                assert w * h / (ai + aj - w * h) < 0.5 + 1e-5


def test_traditional_nms_through_yolact_eval_mode():
    """eval.py --fast_nms=False: net.detect.use_fast_nms = False, then net(x) (eval.py:871)."""
    import yolact_b200
    from oracle.weights import deterministic_state_dict
    g = load_golden("net_resnet50_160")
    cfg = cfg_for(str(g["config"]))
    yolact_b200.cfg.replace(cfg.copy())
    net = yolact_b200.Yolact(cfg, precision="f32")
    net.load_state_dict(deterministic_state_dict(net.state_dict(), int(g["seed"])))
    net.eval()
    net.detect.use_fast_nms = False
    out = net(cuda(g["x"]))
    conf = O.softmax_rows(g["raw_conf"])
    det = O.detect_one(g["raw_loc"][0], conf[0], g["raw_mask"][0], g["raw_priors"], traditional=True,
                       max_size=cfg.max_size)
    got = out[0]["detection"]
    assert got["score"].shape[0] == det["score"].shape[0]
    assert np.array_equal(got["class"].cpu().numpy(), det["class"])
    np.testing.assert_allclose(got["score"].cpu().numpy(), det["score"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(got["box"].cpu().numpy(), det["box"], rtol=0, atol=1e-5)
    net.detect.use_fast_nms = True


# ---- mask_iou / jaccard -----------------------------------------------------------------------------------------
def test_mask_and_box_iou_golden_bit_exact():
    g = load_golden("eval_unit")
    for crowd, tag in ((False, "plain"), (True, "crowd")):
        for dt in (torch.float32, torch.uint8):
            a = mask_iou(cuda(g["iou_masks_a"]).to(dt), cuda(g["iou_masks_b"]).to(dt), crowd).cpu().numpy()
            assert np.array_equal(a, g["iou_mask_" + tag], equal_nan=True)
        flat = mask_iou(cuda(g["iou_masks_a"]).float().view(7, -1), cuda(g["iou_masks_b"]).float().view(5, -1), crowd)
        assert np.array_equal(flat.cpu().numpy(), g["iou_mask_" + tag], equal_nan=True)      # eval.py passes [n, h*w]
        b = jaccard(cuda(g["iou_boxes_a"]), cuda(g["iou_boxes_b"]), crowd).cpu().numpy()
        np.testing.assert_allclose(b, g["iou_box_" + tag], rtol=0, atol=1e-7)


def test_mask_iou_full_size_properties():
    # 100 masks at 550x550 straight from the mask kernel (bit-packed) against themselves and a GT set
    g = load_golden("postprocess_unit")
    r = np.random.RandomState(3)
    n = 100
    coef = np.tanh(r.standard_normal((n, 32))).astype(np.float32)
    c = r.uniform(0.2, 0.8, (n, 2))
    wh = r.uniform(0.1, 0.5, (n, 2))
    box = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    bits, _, _ = assemble_masks(cuda(g["proto"]), cuda(coef), cuda(box), 550, 550, True, "bits")
    u8, _, _ = assemble_masks(cuda(g["proto"]), cuda(coef), cuda(box), 550, 550, True, "u8")
    iou = mask_iou(bits, bits, packed=True).cpu().numpy()
    assert iou.shape == (n, n)
    area = u8.view(n, -1).sum(1).cpu().numpy()
    d = np.diag(iou)
    assert np.all((d == 1) | ((area == 0) & np.isnan(d)))                  # IoU(m, m) == 1 (0/0 for an empty mask)
    assert np.array_equal(iou, iou.T, equal_nan=True)                       # symmetric, bit for bit
    assert np.nanmax(iou) <= 1.0 and np.nanmin(iou) >= 0.0
    # the row-packed layout of the mask kernel and the flat layout of pack_masks agree on the counts
    ref = E.mask_iou(u8[:9].cpu().numpy(), u8[5:20].cpu().numpy())
    got = mask_iou(u8[:9], u8[5:20]).cpu().numpy()
    assert np.array_equal(got, ref, equal_nan=True)
    got2 = mask_iou(bits[:9], bits[5:20], packed=True).cpu().numpy()
    assert np.array_equal(got2, ref, equal_nan=True)
    assert pack_masks(u8[:3]).dtype == torch.int32


# ---- COCO RLE -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["f32", "u8", "bits"])
def test_rle_vs_oracle_small_and_edge_cases(fmt):
    r = np.random.RandomState(4)
    for (h, w) in [(1, 1), (5, 3), (17, 40), (64, 33), (70, 300)]:
        ms = [(r.rand(h, w) < p).astype(np.uint8) for p in (0.0, 0.03, 0.5, 1.0)]
        blob = np.zeros((h, w), np.uint8)
        blob[h // 4:h // 2 + 1, w // 3:w // 2 + 1] = 1
        ms.append(blob)
        m = np.stack(ms)
        if fmt == "bits":
            t = pack_rows(m)
            runs = mask_run_lengths(t, "bits", w=w)
        else:
            t = cuda(m).float() if fmt == "f32" else cuda(m)
            runs = mask_run_lengths(t)
        for i in range(len(ms)):
            assert runs[i].tolist() == E.rle_counts(ms[i]), (h, w, i)


def pack_rows(m):
    """numpy [n,h,w] 0/1 -> the mask kernel's YB_MASK_BITS layout (row pitch ceil(w/32) words) on the GPU."""
    n, h, w = m.shape
    wpr = (w + 31) // 32
    pad = np.zeros((n, h, wpr * 32), np.uint8)
    pad[:, :, :w] = m
    words = (pad.reshape(n, h, wpr, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
    return cuda(words.view(np.int32))


def test_rle_full_size_round_trip_and_overflow_retry():
    g = load_golden("postprocess_unit")
    n = g["coef"].shape[0]
    bits, _, _ = assemble_masks(cuda(g["proto"]), cuda(g["coef"]), cuda(g["box"]), 550, 550, True, "bits")
    u8, _, _ = assemble_masks(cuda(g["proto"]), cuda(g["coef"]), cuda(g["box"]), 550, 550, True, "u8")
    enc = encode_masks(bits, "bits", w=550)
    enc_u8 = encode_masks(u8)
    host = u8.cpu().numpy()
    for i in range(n):
        assert enc[i]["size"] == [550, 550] and enc[i]["counts"] == enc_u8[i]["counts"]
        counts = E.rle_from_string(enc[i]["counts"])
        assert np.array_equal(E.rle_decode(counts, 550, 550), host[i])      # decode(encode(m)) == m
        assert enc[i]["counts"] == E.rle_to_string(E.rle_counts(host[i]))
    # a checkerboard needs h*w runs: the first launch overflows its buffer and the wrapper retries once
    cb = (np.indices((40, 90)).sum(0) % 2).astype(np.uint8)[None]
    runs = mask_run_lengths(cuda(cb), cap=16)
    assert runs[0].tolist() == E.rle_counts(cb[0])


# ---- prep_display -----------------------------------------------------------------------------------------------
def _display_inputs():
    g = load_golden("eval_unit")
    det = {"box": cuda(g["disp_box"]), "mask": cuda(g["disp_coef"]), "class": cuda(g["disp_cls"]),
           "score": cuda(g["disp_score"]), "proto": cuda(g["disp_proto"])}
    return g, det, cuda(g["disp_frame"]).float()


def _caller_prep_display(dets_out, frame, top_k, score_threshold, class_color=False, mask_alpha=0.45):
    """What a caller's prep_display does around the product functions (eval.py:147-167,186-226): postprocess with
    rescore_bbox, argsort + top_k + score cut, palette colours, then ONE display_blend call instead of ~10 ATen ops."""
    cfg = ybcfg.cfg
    save, cfg.rescore_bbox = cfg.rescore_bbox, True                     # eval.py:148-149
    try:
        t = postprocess(dets_out, int(frame.shape[1]), int(frame.shape[0]), crop_masks=True,
                        score_threshold=score_threshold, mask_format="u8")
    finally:
        cfg.rescore_bbox = save
    idx = t[1].argsort(0, descending=True)[:top_k]                      # eval.py:155
    masks = t[3][idx]
    classes, scores = t[0][idx].cpu().numpy(), t[1][idx].cpu().numpy()
    n = min(top_k, classes.shape[0])
    for j in range(n):
        if scores[j] < score_threshold:
            n = j
            break
    colors = [[c / 255.0 for c in get_color(j, classes, class_color, bgr=True)] for j in range(n)]
    return display_blend(frame, masks[:n] if n else None, colors if n else None, mask_alpha).cpu().numpy()


@pytest.mark.parametrize("tag,kw", [("masks", dict(top_k=8, score_threshold=0.15)),
                                    ("classcolor", dict(top_k=15, score_threshold=0.3, class_color=True))])
def test_display_blend_in_a_prep_display_loop_matches_the_reference(tag, kw):
    g, det, frame = _display_inputs()
    ybcfg.set_cfg("yolact_base_config")
    out = _caller_prep_display([{"detection": det, "net": None}], frame, **kw)
    ref = g["disp_" + tag]
    assert out.shape == ref.shape and out.dtype == np.uint8
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 2e-3                      # .byte() truncation: +-1 LSB, rarely


def test_display_blend_formats_and_identity():
    r = np.random.RandomState(9)
    h, w, n = 120, 200, 6
    frame = r.randint(0, 256, size=(h, w, 3)).astype(np.float32)
    m = (r.rand(n, h, w) < 0.3).astype(np.uint8)
    cols = r.uniform(0, 1, (n, 3)).astype(np.float32)
    ref = E.display_blend(frame, m, cols, 0.45)
    for t in (cuda(m), cuda(m).float(), pack_rows(m)):
        out = display_blend(cuda(frame), t, cols, 0.45, w=w).cpu().numpy()
        diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 2e-3
    # no detections: (img / 255 * 255).byte(), exactly the reference's round trip
    out0 = display_blend(cuda(frame), None, None).cpu().numpy()
    assert np.array_equal(out0, E.display_blend(frame, [], [], 0.45))
