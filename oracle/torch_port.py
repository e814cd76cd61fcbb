"""Torch-CPU restatement of Detect (Fast NMS) and postprocess -- TEST INFRASTRUCTURE / CPU baseline.

Same algorithm as oracle/yolact_oracle.py's numpy functions but expressed with the ATen ops the
reference itself runs on a CPU (sort, index, matmul, F.interpolate), so that bench.py's CPU arm times
what the reference would cost on the host cores rather than numpy overhead.  Checked against the
numpy oracle and the golden vectors in tests/test_oracle_vs_golden.py.
Follows layers/functions/detection.py:81-180, layers/box_utils.py:32-80,267-373, layers/output_utils.py:58-99.
"""
import torch
import torch.nn.functional as F


def decode(loc, priors):
    # box_utils.py:303-310
    boxes = torch.cat((priors[:, :2] + loc[:, :2] * 0.1 * priors[:, 2:],
                       priors[:, 2:] * torch.exp(loc[:, 2:] * 0.2)), 1)
    boxes[:, :2] -= boxes[:, 2:] / 2
    boxes[:, 2:] += boxes[:, :2]
    return boxes


def jaccard(a, b):
    # box_utils.py:32-80, [n,A,4] x [n,B,4]
    max_xy = torch.min(a[:, :, None, 2:], b[:, None, :, 2:])
    min_xy = torch.max(a[:, :, None, :2], b[:, None, :, :2])
    inter = torch.clamp(max_xy - min_xy, min=0).prod(3)
    area_a = ((a[:, :, 2] - a[:, :, 0]) * (a[:, :, 3] - a[:, :, 1])).unsqueeze(2)
    area_b = ((b[:, :, 2] - b[:, :, 0]) * (b[:, :, 3] - b[:, :, 1])).unsqueeze(1)
    return inter / (area_a + area_b - inter)


def detect_one(loc, conf, mask, priors, conf_thresh=0.05, nms_thresh=0.5, top_k=200, max_dets=100):
    """conf [P,C] softmaxed.  Stable sorts make the tie order the documented one (lower index first)."""
    boxes = decode(loc, priors)
    cur = conf[:, 1:].t()
    keep = cur.max(dim=0)[0] > conf_thresh
    scores, boxes, masks = cur[:, keep], boxes[keep], mask[keep]
    if scores.size(1) == 0:
        return None
    scores, idx = scores.sort(dim=1, descending=True, stable=True)
    idx, scores = idx[:, :top_k].contiguous(), scores[:, :top_k]
    C, nd = idx.shape
    b = boxes[idx.view(-1)].view(C, nd, 4)
    m = masks[idx.view(-1)].view(C, nd, -1)
    iou = jaccard(b, b).triu_(diagonal=1)
    keepm = iou.max(dim=1)[0] <= nms_thresh
    classes = torch.arange(C)[:, None].expand_as(keepm)[keepm]
    b, m, s = b[keepm], m[keepm], scores[keepm]
    s, o = s.sort(dim=0, descending=True, stable=True)
    o, s = o[:max_dets], s[:max_dets]
    return {"box": b[o], "mask": m[o], "class": classes[o], "score": s}


def sanitize(x1, x2, size, padding=0):
    x1, x2 = x1 * size, x2 * size
    lo, hi = torch.min(x1, x2), torch.max(x1, x2)
    return torch.clamp(lo - padding, min=0), torch.clamp(hi + padding, max=size)


def crop(masks, boxes, padding=1):
    h, w, n = masks.shape
    x1, x2 = sanitize(boxes[:, 0], boxes[:, 2], w, padding)
    y1, y2 = sanitize(boxes[:, 1], boxes[:, 3], h, padding)
    rows = torch.arange(w, dtype=x1.dtype).view(1, -1, 1)
    cols = torch.arange(h, dtype=x1.dtype).view(-1, 1, 1)
    m = (rows >= x1.view(1, 1, -1)) & (rows < x2.view(1, 1, -1)) & (cols >= y1.view(1, 1, -1)) & (cols < y2.view(1, 1, -1))
    return masks * m.float()


def postprocess_one(det, w, h, crop_masks=True, maskiou_fn=None):
    masks = torch.sigmoid(det["proto"] @ det["mask"].t())
    if crop_masks:
        masks = crop(masks, det["box"])
    masks = masks.permute(2, 0, 1).contiguous()
    scores = det["score"]
    if maskiou_fn is not None:
        miou = torch.gather(maskiou_fn(masks.unsqueeze(1)), 1, det["class"].unsqueeze(1)).squeeze(1)
        scores = [scores, scores * miou]
    masks = F.interpolate(masks.unsqueeze(0), (h, w), mode="bilinear", align_corners=False).squeeze(0)
    masks.gt_(0.5)
    x1, x2 = sanitize(det["box"][:, 0], det["box"][:, 2], w)
    y1, y2 = sanitize(det["box"][:, 1], det["box"][:, 3], h)
    return det["class"], scores, torch.stack([x1, y1, x2, y2], 1).long(), masks
