"""Chain kernel (tc_conv.cu: a run of consecutive convolutions -- the bottleneck blocks of a ResNet stage, the protonet's
3x3 stack -- in one persistent launch with per-tile dependency counters) against the same layers launched one by one.

With stream-K off both paths accumulate every output element over the same k-blocks in the same order (the tile shape
does not enter the arithmetic), so the head tensors must be IDENTICAL bit for bit -- any ordering bug between a tile and
the tiles it reads (halo rows, residual) shows up as a difference.  Sizes are chosen so that the chained stages have
many M tiles (dependency ranges cross tile and image boundaries) as well as a single one.
"""
import numpy as np
import pytest
import torch

import yolact_b200
from oracle.weights import deterministic_state_dict, deterministic_input
from tests.helpers import cfg_for

pytestmark = pytest.mark.gpu

CASES = [
    # config, batch, height, width
    ("yolact_base_config", 2, 550, 550),          # the whole trunk after the max-pool (102 layers, 4 resolutions), FPN, protonet
    ("yolact_base_config", 3, 256, 320),          # non-square, odd batch: flat tiles straddle rows and images
    ("yolact_resnet50_config", 1, 160, 160),      # a single M tile per layer
    ("yolact_im700_config", 1, 700, 700),
    ("yolact_plus_resnet50_config", 2, 256, 256), # DCN blocks cut the trunk into several runs
    ("yolact_darknet53_config", 2, 320, 320),     # 3x3 + residual layers are not chainable: short runs only
]


def _run(cfg_name, B, H, W, chain, monkeypatch, precision="f16x3"):
    monkeypatch.setenv("YB_CHAIN", str(chain))
    monkeypatch.setenv("YB_SK", "0")
    cfg = cfg_for(cfg_name)
    yolact_b200.cfg.replace(cfg.copy())
    net = yolact_b200.Yolact(cfg, precision=precision)
    net.load_state_dict(deterministic_state_dict(net.state_dict(), 7))
    net.train()
    x = deterministic_input(B, H, W, seed=11).cuda()
    out = net(x)
    out = {k: out[k].cpu().numpy() for k in ("proto", "loc", "conf", "mask")}
    again = net(x)                                  # the dependency counters are re-armed by every launch
    for k in out:
        assert np.array_equal(out[k], again[k].cpu().numpy()), k
    names = [n for n, _ in net.profile_conv_stack(x)]
    del net
    torch.cuda.empty_cache()
    return out, names


@pytest.mark.parametrize("precision", ["f16x3", "f16tc"])
@pytest.mark.parametrize("cfg_name,B,H,W", CASES)
def test_chain_identical_to_separate_launches(cfg_name, B, H, W, precision, monkeypatch):
    sep, names_sep = _run(cfg_name, B, H, W, 0, monkeypatch, precision)
    ch, names_ch = _run(cfg_name, B, H, W, 2, monkeypatch, precision)
    assert not any(n.startswith("chain") for n in names_sep)
    chains = [n for n in names_ch if n.startswith("chain")]
    assert chains, "YB_CHAIN=2 formed no chain"
    print(cfg_name, B, H, W, precision, "chains:", [c.split(" [")[0] for c in chains], "ops", len(names_sep), "->", len(names_ch))
    for k in sep:
        assert np.isfinite(ch[k]).all()
        assert np.array_equal(sep[k], ch[k]), (k, float(np.abs(sep[k] - ch[k]).max()))
