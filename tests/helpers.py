"""Shared test helpers (config objects, comparison utilities)."""
import numpy as np

from yolact_b200.config import CONFIGS


def cfg_for(name):
    return CONFIGS[name].copy()


def unpack_masks(packed, w):
    return np.unpackbits(packed, axis=-1)[..., :w].astype(np.float32)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))
