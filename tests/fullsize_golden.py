"""Helpers for the FULL-SIZE reference goldens (tests/golden/full_*.npz, oracle/gen_golden.py `full`): one image of a
BASELINE.json config run through the REAL reference at full resolution (raw heads subsampled, detections complete)."""
import numpy as np

from tests.conftest import load_golden
from tests.helpers import cfg_for, unpack_masks

FULL_CASES = ["full_base_550", "full_base_550_to_480x640", "full_plus_resnet50_550", "full_im700_700",
              "full_plus_base_550"]


def load_case(tag):
    g = load_golden(tag)
    cfg = cfg_for(str(g["config"]))
    ph, pw = (int(v) for v in g["post_hw"])
    ref = {"class": g["det_class"], "score": g["det_score"], "box": g["det_box"], "coef": g["det_mask"],
           "box_px": g["post_boxes"], "masks": unpack_masks(g["post_masks_packed"], pw) > 0.5,
           "score_maskiou": g["post_scores_maskiou"] if "post_scores_maskiou" in g.files else None}
    return g, cfg, ref, (ph, pw)


def raw_errors(raw, g):
    """max |got - golden| / max |full golden tensor| of the subsampled head tensors ([1, ...] batch)."""
    hs, ps = int(g["head_stride"]), int(g["proto_stride"])
    out = {}
    for k in ("loc", "conf", "mask"):
        got = np.asarray(raw[k], np.float64)[:1, ::hs]
        out[k] = float(np.abs(got - g["raw_" + k]).max() / float(g["raw_%s_absmax" % k]))
    got = np.asarray(raw["proto"], np.float64)[:1, ::ps, ::ps]
    out["proto"] = float(np.abs(got - g["raw_proto"]).max() / float(g["raw_proto_absmax"]))
    out["priors_equal"] = bool(np.array_equal(np.asarray(raw["priors"]), g["raw_priors"]))
    return out
