"""The reference evaluation driver's call sequences, end to end on the GPU through the product functions only, against
the CPU oracle: what `eval.py --benchmark` (prep_benchmark, eval.py:264-281) and `eval.py` mAP mode (prep_metrics,
eval.py:385-445, incl. the --output_coco_json branch :417-427) do with one network output.  The bookkeeping around the
calls (APDataObject, Detections) is caller code and is not rebuilt; every tensor it consumes is checked here."""
import numpy as np
import pytest
import torch

import yolact_b200
from oracle import eval_oracle as E
from oracle import torch_port as T
from oracle import yolact_oracle as O
from oracle.weights import deterministic_input, deterministic_state_dict
from tests.helpers import cfg_for
from tests.parity_utils import align
from yolact_b200.eval_utils import encode_masks, jaccard, mask_iou
from yolact_b200.output_utils import postprocess

pytestmark = pytest.mark.gpu

H, W = 200, 264          # "original image" size the masks are produced at (eval.py:266,403: postprocess(dets, w, h))
_state = {}


def _setup():
    if _state:
        return _state
    cfg = cfg_for("yolact_resnet50_config")
    yolact_b200.cfg.replace(cfg.copy())
    net = yolact_b200.Yolact(cfg)                      # default precision: the benched split-precision mode
    net.detect.use_fast_nms = True                     # eval.py:871
    sd = deterministic_state_dict(net.state_dict(), 0)
    net.load_state_dict(sd)
    net.eval()
    x = deterministic_input(1, 256, 256, 321)
    preds = net(x.cuda())                              # eval.py:945
    # the oracle's view of the same image
    orc = O.ConvStackOracle(cfg, sd)
    raw = orc.forward(x)
    det = T.detect_one(raw["loc"][0], torch.softmax(raw["conf"], -1)[0], raw["mask"][0], raw["priors"])
    det["proto"] = raw["proto"][0]
    box_rel = det["box"].clone().numpy()
    oc, osc, obx, oms = T.postprocess_one(det, W, H)
    ref = {"class": oc.numpy(), "score": osc.numpy(), "box": box_rel, "box_px": obx.numpy(), "masks": oms.numpy()}
    _state.update(cfg=cfg, net=net, preds=preds, ref=ref)
    return _state


def _aligned(st):
    """postprocess output re-ordered to the oracle's ranking (ranks may swap only between ties, parity_utils.align)."""
    d = st["preds"][0]["detection"]
    got = {"class": d["class"].cpu().numpy(), "score": d["score"].cpu().numpy(), "box": d["box"].cpu().numpy()}
    perm, ok = align(got, st["ref"])
    assert ok
    return perm


def test_prep_benchmark_sequence():
    st = _setup()
    top_k = 5                                                                        # eval.py:46
    t = postprocess(st["preds"], W, H, crop_masks=True, score_threshold=0)           # eval.py:266
    classes, scores, boxes, masks = [x[:top_k] for x in t]                           # eval.py:269
    classes, scores, boxes, masks = classes.cpu().numpy(), scores.cpu().numpy(), boxes.cpu().numpy(), masks.cpu().numpy()
    torch.cuda.synchronize()                                                         # eval.py:281
    perm, ref = _aligned(st), st["ref"]
    assert classes.dtype == np.int64 and masks.dtype == np.float32 and masks.shape == (top_k, H, W)
    full = [x.cpu().numpy() for x in t]
    assert np.array_equal(full[0][perm][:top_k], ref["class"][:top_k])
    np.testing.assert_allclose(full[1][perm][:top_k], ref["score"][:top_k], atol=1e-3)
    assert np.abs(full[2][perm][:top_k] - ref["box_px"][:top_k]).max() <= 1
    assert (full[3][perm][:top_k] != ref["masks"][:top_k]).mean() < 1e-3


def test_prep_metrics_sequence():
    st = _setup()
    ref = st["ref"]
    r = np.random.RandomState(5)
    n_gt, num_crowd = 7, 2
    gt = np.zeros((n_gt, 5), np.float32)                                             # [x1, y1, x2, y2, class], relative
    gt[:, :2] = r.uniform(0.0, 0.6, (n_gt, 2))
    gt[:, 2:4] = gt[:, :2] + r.uniform(0.1, 0.4, (n_gt, 2))
    gt[:, 4] = r.randint(0, 80, n_gt)
    gt_masks_np = (r.rand(n_gt, H, W) < 0.3).astype(np.uint8)
    # eval.py:389-401 ("Prepare gt")
    gt_boxes = torch.Tensor(gt[:, :4])
    gt_boxes[:, [0, 2]] *= W
    gt_boxes[:, [1, 3]] *= H
    gt_masks = torch.Tensor(gt_masks_np).view(-1, H * W)
    split = lambda x: (x[-num_crowd:], x[:-num_crowd])
    crowd_boxes, gt_boxes = split(gt_boxes)
    crowd_masks, gt_masks = split(gt_masks)
    # eval.py:403-416 ("Postprocess")
    classes, scores, boxes, masks = postprocess(st["preds"], W, H, crop_masks=True, score_threshold=0)
    assert classes.size(0) > 0
    masks = masks.view(-1, H * W).cuda()
    boxes = boxes.cuda()
    # eval.py:433-441 ("Eval Setup"): the four IoU caches
    mask_iou_cache = mask_iou(masks, gt_masks.cuda()).cpu().numpy()
    bbox_iou_cache = jaccard(boxes.float(), gt_boxes.float().cuda()).cpu().numpy()
    crowd_mask_iou_cache = mask_iou(masks, crowd_masks.cuda(), iscrowd=True).cpu().numpy()
    crowd_bbox_iou_cache = jaccard(boxes.float(), crowd_boxes.float().cuda(), iscrowd=True).cpu().numpy()
    # the oracle on ITS detections; rows compared through the tie-aware alignment
    perm = _aligned(st)
    om = ref["masks"].reshape(len(ref["class"]), -1)
    ob = ref["box_px"].astype(np.float32)
    same_masks = (masks.cpu().numpy()[perm] == om).all(axis=1)                      # rows whose masks are pixel-identical
    assert same_masks.mean() > 0.9
    want = E.mask_iou(om, gt_masks.numpy(), False)
    assert np.array_equal(mask_iou_cache[perm][same_masks], want[same_masks], equal_nan=True)   # integer counts: exact
    want = E.mask_iou(om, crowd_masks.numpy(), True)
    assert np.array_equal(crowd_mask_iou_cache[perm][same_masks], want[same_masks], equal_nan=True)
    same_boxes = (boxes.cpu().numpy()[perm] == ref["box_px"]).all(axis=1)
    assert same_boxes.mean() > 0.9
    np.testing.assert_allclose(bbox_iou_cache[perm][same_boxes], E.box_iou(ob, gt_boxes.numpy(), False)[same_boxes], atol=1e-6)
    np.testing.assert_allclose(crowd_bbox_iou_cache[perm][same_boxes], E.box_iou(ob, crowd_boxes.numpy(), True)[same_boxes], atol=1e-6)


def test_output_coco_json_sequence():
    """eval.py:417-427 -> Detections.add_mask (eval.py:320-330): pycocotools.mask.encode of every kept mask."""
    st = _setup()
    classes, scores, boxes, masks = postprocess(st["preds"], W, H, crop_masks=True, score_threshold=0)
    rles = encode_masks(masks)                                                       # one launch + one small D2H
    m = masks.cpu().numpy()
    assert len(rles) == m.shape[0]
    for i in range(0, m.shape[0], 7):
        assert rles[i]["size"] == [H, W]
        counts = E.rle_counts(m[i].astype(np.uint8))
        assert rles[i]["counts"] == E.rle_to_string(counts)
        assert np.array_equal(E.rle_decode(E.rle_from_string(rles[i]["counts"]), H, W), m[i].astype(np.uint8))
