"""Deterministic synthetic weights shared by oracle/gen_golden.py (applied to the REFERENCE model in
the build container) and by the tests (applied to yolact_b200.Yolact / the oracle on the GPU box).

There are no checkpoints in the container, and default-initialised weights produce zero detections
(softmax conf ~ 1/81 < 0.05).  These weights are a function of (key name,
shape, seed) only, so both sides can rebuild them bit-identically without shipping ~200 MB:
  * conv weights ~ N(0, gain^2 * 2 / fan_in)  (activations stay O(1) through 101 layers)
  * BatchNorm: gamma in [0.8,1.2] (x0.4 for the last BN of a bottleneck), beta in [-0.1,0.1],
    running_mean in [-0.2,0.2], running_var in [0.6,1.4]  -> BN folding is actually exercised
  * conf_layer scaled so that a few hundred priors clear the 0.05 score threshold with well
    separated scores; conv_offset_mask non-zero so DCN really deforms (the reference zero-inits it)
"""
import zlib

import numpy as np
import torch


def _rng(name, seed):
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)


def deterministic_tensor(name, shape, seed=0, is_bn=False):
    r = _rng(name, seed)
    shape = tuple(int(s) for s in shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return torch.from_numpy(r.uniform(-0.2, 0.2, shape).astype(np.float32))
    if leaf == "running_var":
        return torch.from_numpy(r.uniform(0.6, 1.4, shape).astype(np.float32))
    if is_bn and leaf == "weight":
        g = r.uniform(0.8, 1.2, shape).astype(np.float32)
        if ".bn3." in name or ".conv2.1." in name:
            g *= 0.2     # last BN of a residual block: keeps the residual stream from growing with depth
        return torch.from_numpy(g)
    if is_bn and leaf == "bias":
        return torch.from_numpy(r.uniform(-0.1, 0.1, shape).astype(np.float32))
    if leaf == "weight" and len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        gain = 1.0
        if "conf_layer" in name:
            gain = 0.6
        elif "bbox_layer" in name:
            gain = 0.5
        elif "fpn.lat_layers" in name:
            gain = 0.3
        elif "mask_layer" in name:
            gain = 0.35
        elif "conv_offset_mask" in name:
            gain = 0.5
        elif "maskiou_net" in name:
            gain = 1.4
        w = r.standard_normal(shape).astype(np.float32) * np.float32(gain * np.sqrt(2.0 / fan_in))
        return torch.from_numpy(w)
    if leaf == "bias":
        b = r.uniform(-0.05, 0.05, shape).astype(np.float32)
        if "conf_layer" in name:
            # background logit pushed up so most priors are background, like a trained net
            b = b.reshape(-1, 81) if b.size % 81 == 0 else b
            if b.ndim == 2:
                b[:, 0] += 9.0
            b = b.reshape(shape)
        return torch.from_numpy(np.ascontiguousarray(b))
    return torch.from_numpy(r.standard_normal(shape).astype(np.float32) * 0.05)


def deterministic_state_dict(template, seed=0):
    """template: a state_dict (or {name: tensor-like with .shape}); returns a new dict with the same keys."""
    bn = {k[:-len(".running_mean")] for k in template if k.endswith(".running_mean")}
    return {k: deterministic_tensor(k, v.shape, seed, is_bn=k.rsplit(".", 1)[0] in bn) for k, v in template.items()}


def deterministic_input(B, H, W, seed=1234):
    """N(0,1) frames standing in for (img - MEANS) / STD (data/config.py:28-29)."""
    r = np.random.RandomState(seed)
    return torch.from_numpy(r.standard_normal((B, 3, H, W)).astype(np.float32))
