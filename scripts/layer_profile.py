"""Per-layer device times of the conv stack (eager, CUDA events per op) -> markdown on stdout."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yolact_b200
from oracle.weights import deterministic_state_dict, deterministic_input
from yolact_b200.config import CONFIGS
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolact_base_config"); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--precision", default="f16x3")
a = ap.parse_args()
cfg = CONFIGS[a.config].copy(); yolact_b200.cfg.replace(cfg.copy())
net = yolact_b200.Yolact(cfg, precision=a.precision); net.load_state_dict(deterministic_state_dict(net.state_dict(), 0)); net.eval(); net.detect.use_fast_nms = True
x = deterministic_input(a.batch, cfg.max_size, cfg.max_size, 1).cuda()
net.profile_conv_stack(x)
prof = net.profile_conv_stack(x)
tot = sum(ms for _, ms in prof)
print("# per-layer conv-stack times, %s batch %d, precision %s (eager, CUDA events, warm L2): total %.3f ms\n" % (cfg.name, a.batch, a.precision, tot))
print("| layer | ms | GFLOP | TFLOP/s |\n|---|---:|---:|---:|")
import re
for name, ms in prof:
    m = re.search(r"(\d+)->(\d+) k(\d+)s(\d+) (\d+)x(\d+)", name)
    gf = 0.0
    if m:
        ci, co, k, s, ho, wo = map(int, m.groups())
        gf = 2.0 * a.batch * ho * wo * ci * co * k * k / 1e9
    m = re.search(r"gflop=([0-9.]+)", name)      # a chain of layers in one launch carries its total
    if m:
        gf = float(m.group(1))
    print("| %s | %.4f | %.2f | %.0f |" % (name, ms, gf, gf / ms if ms > 0 else 0))
