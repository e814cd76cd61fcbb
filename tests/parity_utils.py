"""End-to-end parity measurement of a yolact_b200.Yolact against the CPU oracle (TEST INFRASTRUCTURE).

`measure(cfg, precision, batch, size, out_hw)` runs `net(x)` in eval mode + `postprocess()` per image on
the GPU -- the call sequence of eval.py (eval.py:949, :264-281) -- and the same images through the CPU
oracle (oracle/yolact_oracle.py conv stack + oracle/torch_port.py Detect/postprocess), and returns the
quantities north_star's tolerance is stated on: max |dbox|, max |dscore|, class-id equality, NMS keep-set
agreement and the fraction of flipped mask pixels, plus the raw head tensors' error.
"""
import numpy as np
import torch


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def oracle_outputs(cfg, sd, x, out_hw):
    """CPU oracle: raw head tensors + per-image (classes, scores, boxes_rel, boxes_px, masks)."""
    from oracle import yolact_oracle as O
    from oracle import torch_port as T
    orc = O.ConvStackOracle(cfg, sd)
    raw = orc.forward(x)
    conf = torch.softmax(raw["conf"], -1)
    per_image = []
    with torch.no_grad():
        for b in range(x.shape[0]):
            det = T.detect_one(raw["loc"][b], conf[b], raw["mask"][b], raw["priors"], cfg.nms_conf_thresh,
                               cfg.nms_thresh, cfg.nms_top_k, cfg.max_num_detections)
            if det is None:
                per_image.append(None)
                continue
            det["proto"] = raw["proto"][b]
            fn = orc.maskiou if cfg.use_maskiou else None
            box_rel = det["box"].clone()
            classes, scores, boxes_px, masks = T.postprocess_one(det, out_hw[1], out_hw[0], maskiou_fn=fn)
            s2 = None
            if isinstance(scores, list):
                scores, s2 = scores
            per_image.append({"class": classes.numpy(), "score": scores.numpy(), "box": box_rel.numpy(),
                              "box_px": boxes_px.numpy(), "masks": masks.numpy() > 0.5,
                              "score_maskiou": None if s2 is None else s2.numpy()})
    return {k: v.numpy() for k, v in raw.items()}, per_image


def gpu_outputs(net, x, out_hw):
    import yolact_b200
    from yolact_b200.output_utils import postprocess
    dev = torch.device("cuda", torch.cuda.current_device())
    net.train()
    raw = {k: v.cpu().numpy() for k, v in net(x.to(dev)).items()}
    net.eval()
    preds = net(x.to(dev))
    per_image = []
    for b, p in enumerate(preds):
        det = p["detection"]
        if det is None:
            per_image.append(None)
            continue
        box_rel = det["box"].cpu().numpy().copy()
        classes, scores, boxes_px, masks = postprocess(preds, out_hw[1], out_hw[0], batch_idx=b)
        s2 = None
        if isinstance(scores, list):
            scores, s2 = scores
        per_image.append({"class": classes.cpu().numpy(), "score": scores.cpu().numpy(), "box": box_rel,
                          "box_px": boxes_px.cpu().numpy(), "masks": masks.cpu().numpy() > 0.5,
                          "score_maskiou": None if s2 is None else s2.cpu().numpy()})
    return raw, per_image


TIE_EPS = 2e-5   # scores closer than this are ties: fp32 summation order decides their rank in ANY implementation


def align(g, r, tie_eps=TIE_EPS):
    """Pairs every reference detection with a GPU detection of the same class (nearest box).  Returns (perm, ok):
    perm[i] = GPU row matched to reference row i (or -1), ok = the pairing is a permutation that only moves rows
    across reference scores closer than tie_eps (i.e. equal ranking up to numerical ties)."""
    n = len(r["score"])
    used = np.zeros(len(g["score"]), bool)
    perm = -np.ones(n, np.int64)
    for i in range(n):
        cand = np.nonzero((g["class"] == r["class"][i]) & ~used)[0]
        if cand.size == 0:
            continue
        d = np.abs(g["box"][cand] - r["box"][i]).max(axis=1)
        k = int(d.argmin())
        if d[k] < 1e-3:
            perm[i] = cand[k]
            used[cand[k]] = True
    ok = len(g["score"]) == n and bool((perm >= 0).all())
    if ok:
        for i in range(n):
            jdx = int(perm[i])
            if jdx != i and abs(float(r["score"][i]) - float(r["score"][jdx])) >= tie_eps:
                ok = False
                break
    return perm, ok


def compare(raw_g, img_g, raw_r, img_r):
    """Returns a flat dict of parity figures (worst case over the batch)."""
    out = {}
    for k in ("loc", "conf", "mask", "proto"):
        out["raw_" + k] = _rel(raw_g[k], raw_r[k])
    out["priors_equal"] = bool(np.array_equal(raw_g["priors"], raw_r["priors"]))
    n_img = len(img_r)
    strict, modties, keep_agree, dbox, dscore, dboxpx, flips, counts = [], [], [], [], [], [], [], []
    dmiou = []
    for g, r in zip(img_g, img_r):
        if r is None or g is None:
            strict.append(g is None and r is None)
            modties.append(g is None and r is None)
            continue
        counts.append((len(g["score"]), len(r["score"])))
        strict.append(bool(len(g["class"]) == len(r["class"]) and np.array_equal(g["class"], r["class"])))
        perm, ok = align(g, r)
        modties.append(bool(ok))
        keep_agree.append(float((perm >= 0).mean()))       # reference detections found in the GPU's keep set
        m = perm >= 0
        if m.any():
            pg = perm[m]
            dbox.append(float(np.abs(g["box"][pg] - r["box"][m]).max()))
            dscore.append(float(np.abs(g["score"][pg] - r["score"][m]).max()))
            dboxpx.append(int(np.abs(g["box_px"][pg] - r["box_px"][m]).max()))
            flips.append(float((g["masks"][pg] != r["masks"][m]).mean()))
            if r["score_maskiou"] is not None:
                dmiou.append(float(np.abs(g["score_maskiou"][pg] - r["score_maskiou"][m]).max()))
    out["images"] = n_img
    out["counts_gpu_ref"] = counts
    out["class_ids_equal_strict"] = bool(all(strict))            # same class at every rank
    out["class_ids_equal_strict_images"] = int(sum(strict))
    out["class_ids_equal"] = bool(all(modties))                  # ... up to swaps between scores closer than TIE_EPS
    out["class_ids_equal_images"] = int(sum(modties))
    out["keep_set_agreement_min"] = float(min(keep_agree)) if keep_agree else None
    out["max_abs_dbox"] = max(dbox) if dbox else None
    out["max_abs_dscore"] = max(dscore) if dscore else None
    out["max_abs_dbox_px"] = max(dboxpx) if dboxpx else None
    out["mask_pixel_mismatch_max"] = max(flips) if flips else None
    out["max_abs_dscore_maskiou"] = max(dmiou) if dmiou else None
    return out


def measure(cfg, precision, batch, size, out_hw=None, seed_w=0, seed_x=99, net=None):
    import yolact_b200
    from oracle.weights import deterministic_state_dict, deterministic_input
    out_hw = out_hw or (size, size)
    yolact_b200.cfg.replace(cfg.copy())
    own = net is None
    if own:
        net = yolact_b200.Yolact(cfg, precision=precision)
        net.detect.use_fast_nms = True   # what eval.py does from --fast_nms (default True, eval.py:50,871)
        sd = deterministic_state_dict(net.state_dict(), seed_w)
        net.load_state_dict(sd)
    else:
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    x = deterministic_input(batch, size, size, seed_x)
    raw_g, img_g = gpu_outputs(net, x, out_hw)
    if own:
        del net
        torch.cuda.empty_cache()
    raw_r, img_r = oracle_outputs(cfg, sd, x, out_hw)
    r = compare(raw_g, img_g, raw_r, img_r)
    r.update({"config": cfg.name, "precision": precision, "batch": batch, "size": size, "out_hw": list(out_hw)})
    return r
