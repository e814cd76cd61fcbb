"""CPU oracle for the rows either side of the hot path -- TEST INFRASTRUCTURE,
never imported by the product.

  fast_base_transform : FastBaseTransform.forward (utils/augmentations.py:616-658)
  mask_iou / box_iou  : layers/box_utils.py:98-113 / :54-79 (eval.py:435-445 are the callers)
  rle_*               : COCO run-length encoding.  The reference calls pycocotools.mask.encode
                        (eval.py:320-330); pycocotools (requirement of the reference, version unpinned in its
                        README / environment.yml, 2.0.x at the time) is NOT vendored in /root/reference and not
                        installed here, so rleEncode / rleToString / rleFrString of its common/maskApi.c are
                        restated from the published algorithm.  PARITY UNPINNED for the compressed string
                        (no pycocotools to generate vectors with); the run lengths themselves are pinned by
                        the decode(encode(m)) == m round trip and a hand-checked vector in the tests.
  display_blend       : the mask blend of prep_display (eval.py:186-209) + `(img_gpu * 255).byte()` (:226)

Pinned against the real reference by tests/golden/eval_unit.npz (oracle/gen_golden.py).
"""
import numpy as np

from oracle.yolact_oracle import bilinear_resize

MEANS = (103.94, 116.78, 123.68)   # data/config.py:28 (BGR)
STD = (57.38, 57.12, 58.40)        # data/config.py:29


def fast_base_transform(img, out_h, out_w, mode="normalize", mean=MEANS, std=STD):
    """img [B,H,W,3] BGR (uint8 or float) -> [B,3,out_h,out_w] float32 RGB."""
    x = np.asarray(img).astype(np.float32).transpose(0, 3, 1, 2)            # :637 permute
    B, C, H, W = x.shape
    x = bilinear_resize(x.reshape(B * C, H, W), out_h, out_w).reshape(B, C, out_h, out_w)   # :638
    mean = np.asarray(mean, np.float32).reshape(1, 3, 1, 1)
    std = np.asarray(std, np.float32).reshape(1, 3, 1, 1)
    if mode == "normalize":
        x = ((x - mean) / std).astype(np.float32)                          # :641
    elif mode == "subtract_means":
        x = (x - mean).astype(np.float32)
    elif mode == "to_float":
        x = (x / np.float32(255)).astype(np.float32)
    return np.ascontiguousarray(x[:, ::-1])                                # :651  BGR -> RGB


def mask_iou(a, b, iscrowd=False):
    """a [n,h,w], b [m,h,w] 0/1 -> [n,m] float32 (box_utils.py:98-113)."""
    a = np.asarray(a, np.float32).reshape(a.shape[0], -1)
    b = np.asarray(b, np.float32).reshape(b.shape[0], -1)
    inter = (a.astype(np.float64) @ b.astype(np.float64).T).astype(np.float32)   # exact integers
    area_a = a.sum(1, dtype=np.float64).astype(np.float32)[:, None]
    area_b = b.sum(1, dtype=np.float64).astype(np.float32)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        if iscrowd:
            return (inter / area_a).astype(np.float32)
        return (inter / ((area_a + area_b).astype(np.float32) - inter).astype(np.float32)).astype(np.float32)


def box_iou(a, b, iscrowd=False):
    """jaccard, box_utils.py:54-79: a [n,4], b [m,4]."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    max_xy = np.minimum(a[:, None, 2:], b[None, :, 2:])
    min_xy = np.maximum(a[:, None, :2], b[None, :, :2])
    wh = np.clip((max_xy - min_xy).astype(np.float32), 0, None)
    inter = (wh[..., 0] * wh[..., 1]).astype(np.float32)
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])).astype(np.float32)[:, None]
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(np.float32)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        if iscrowd:
            return (inter / area_a).astype(np.float32)
        return (inter / ((area_a + area_b).astype(np.float32) - inter).astype(np.float32)).astype(np.float32)


# ---- COCO RLE (pycocotools common/maskApi.c: rleEncode, rleToString, rleFrString, rleDecode) -----------------
def rle_counts(mask):
    """mask [h,w] 0/1 -> list of run lengths over the column-major flattening, starting with the zeros run."""
    m = np.asarray(mask).astype(bool).T.reshape(-1)      # Fortran order
    if m.size == 0:
        return [0]
    change = np.flatnonzero(m[1:] != m[:-1]) + 1
    pos = np.concatenate([[0], change, [m.size]])
    runs = np.diff(pos).tolist()
    if m[0]:
        runs = [0] + runs
    return runs


def rle_to_string(counts):
    """rleToString: LEB128-like, 5 data bits per char, deltas against counts[i-2] for i > 2."""
    out = bytearray()
    for i, x in enumerate(counts):
        x = int(x)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1f
            x >>= 5                                        # arithmetic shift on negative ints, like C's long
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def rle_from_string(s):
    """rleFrString."""
    counts = []
    p = 0
    s = bytes(s)
    while p < len(s):
        x = 0
        k = 0
        more = True
        while more:
            c = s[p] - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def rle_decode(counts, h, w):
    v = np.zeros(h * w, np.uint8)
    p = 0
    val = 0
    for c in counts:
        v[p:p + c] = val
        p += c
        val ^= 1
    return np.ascontiguousarray(v.reshape(w, h).T)


def display_blend(img, masks, colors, alpha, img_is_255=True):
    """img [h,w,3] float; masks [n,h,w] 0/1 in drawing order; colors [n,3] in 0..1 -> uint8 [h,w,3]
    (eval.py:186-209 then :226).  Sequential form `img = img * inv_alph[j] + masks_color[j]` is algebraically
    what the reference's cumprod expression evaluates; this restatement follows the reference's expression."""
    v = np.asarray(img, np.float32)
    if img_is_255:
        v = (v / np.float32(255.0)).astype(np.float32)                       # :144
    n = len(masks)
    if n > 0:
        m = np.asarray(masks, np.float32)[:, :, :, None]
        col = np.asarray(colors, np.float32).reshape(n, 1, 1, 3)
        masks_color = ((np.repeat(m, 3, axis=3) * col).astype(np.float32) * np.float32(alpha)).astype(np.float32)
        inv = (m * np.float32(-alpha) + np.float32(1)).astype(np.float32)
        summand = masks_color[0].copy()
        if n > 1:
            cum = np.cumprod(inv[:n - 1], axis=0, dtype=np.float32)
            summand = summand + (masks_color[1:] * cum).astype(np.float32).sum(axis=0, dtype=np.float32)
        v = (v * np.prod(inv, axis=0, dtype=np.float32) + summand).astype(np.float32)
    return np.clip((v * np.float32(255)).astype(np.float32), 0, 255).astype(np.uint8)   # .byte(): truncation


COLORS = ((244, 67, 54), (233, 30, 99), (156, 39, 176), (103, 58, 183), (63, 81, 181), (33, 150, 243), (3, 169, 244),
          (0, 188, 212), (0, 150, 136), (76, 175, 80), (139, 195, 74), (205, 220, 57), (255, 235, 59), (255, 193, 7),
          (255, 152, 0), (255, 87, 34), (121, 85, 72), (158, 158, 158), (96, 125, 139))   # data/config.py:6-24


def prep_display_masks(det, frame, top_k=5, score_threshold=0.0, class_color=False, mask_alpha=0.45, crop=True):
    """prep_display (eval.py:135-226) with undo_transform=False and text / boxes off: postprocess at the frame size
    with rescore_bbox, score_threshold filter, top_k by score, colour j*5 (or class*5) mod 19 in BGR, blend."""
    from oracle.yolact_oracle import postprocess_one, _stable_desc_order
    h, w = frame.shape[:2]
    det = {k: np.asarray(v) for k, v in det.items()}
    if score_threshold > 0:                                          # output_utils.py:42-50
        keep = det["score"] > np.float32(score_threshold)
        det = {k: (v if k == "proto" else v[keep]) for k, v in det.items()}
    if det["score"].shape[0] == 0:
        return display_blend(frame, [], [], mask_alpha)
    classes, scores, boxes, masks = postprocess_one(det, w, h, crop_masks=crop)
    idx = _stable_desc_order(np.asarray(scores))[:top_k]             # eval.py:156
    classes, scores, masks = classes[idx], np.asarray(scores)[idx], masks[idx]
    n = min(top_k, classes.shape[0])
    for j in range(n):
        if scores[j] < score_threshold:
            n = j
            break
    cols = []
    for j in range(n):
        c = COLORS[(int(classes[j]) * 5 if class_color else j * 5) % len(COLORS)]
        cols.append([np.float32(c[2]) / np.float32(255.), np.float32(c[1]) / np.float32(255.), np.float32(c[0]) / np.float32(255.)])
    return display_blend(frame, masks[:n], cols, mask_alpha)
