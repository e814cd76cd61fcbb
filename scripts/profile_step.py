"""One steady-state step of the bench workload between cudaProfilerStart/Stop, for
   ncu --profile-from-start off ... python scripts/profile_step.py [--config ...] [--batch 8] [--eager]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolact_b200
from oracle.weights import deterministic_state_dict, deterministic_input
from yolact_b200.config import CONFIGS
from yolact_b200.output_utils import assemble_masks_batch

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolact_base_config")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--conv-only", action="store_true")
args = ap.parse_args()
cfg = CONFIGS[args.config].copy()
size = args.size or cfg.max_size
yolact_b200.cfg.replace(cfg.copy())
net = yolact_b200.Yolact(cfg, precision=os.environ.get("YB_PRECISION", "f16x3")); net.detect.use_fast_nms = True
net.load_state_dict(deterministic_state_dict(net.state_dict(), 0))
net.eval()
x = deterministic_input(args.batch, size, size, 1234).cuda()
masks = torch.empty(args.batch, 100, size, size, device="cuda")


def step():
    if args.conv_only:
        net.forward_conv_only(x)
        return
    box, coef, cls, score, count, proto = net.infer_padded(x)
    assemble_masks_batch(proto, coef, box, size, size, True, "f32", masks_out=masks)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.steps, "step(s)")
