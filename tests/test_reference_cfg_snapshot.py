"""from_reference_cfg(): snapshotting the REAL reference cfg gives our named configs.  Runs only where
/root/reference exists (the build container); skipped on the GPU box."""
import os
import sys
import types

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_snapshot_matches_named_configs():
    import torch
    for m in ["pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval", "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["pycocotools.coco"].COCO = object
    sys.path.insert(0, REF)
    try:
        from data import config as rc
        from yolact_b200.config import CONFIGS, from_reference_cfg
        for name, mine in CONFIGS.items():
            snap = from_reference_cfg(getattr(rc, name))
            for key in ("backbone", "backbone_layers", "dcn_layers", "dcn_interval", "selected_layers", "max_size",
                        "pred_aspect_ratios", "use_square_anchors", "num_classes", "fpn_features", "use_maskiou",
                        "rescore_mask", "rescore_bbox", "nms_top_k", "nms_conf_thresh", "nms_thresh",
                        "max_num_detections", "normalize", "to_float"):
                assert getattr(snap, key) == getattr(mine, key), (name, key)
            for a, b in zip(snap.pred_scales, mine.pred_scales):
                assert [float(x) for x in a] == [float(x) for x in b], name
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "data" or k.startswith("data.") or k in ("backbone",)]:
            del sys.modules[k]
