"""Device time of the batched mask-assembly kernel for a (size, batch) and every format."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolact_b200.output_utils import assemble_masks_batch
size, B = int(sys.argv[1]), int(sys.argv[2])
ps = {550: 138, 700: 176}.get(size, size // 4)
r = np.random.RandomState(0)
n = 100
proto = torch.from_numpy(np.maximum(r.standard_normal((B, ps, ps, 32)), 0).astype(np.float32)).cuda()
coef = torch.from_numpy(np.tanh(r.standard_normal((B, n, 32))).astype(np.float32)).cuda()
c = r.uniform(0.2, 0.8, (B, n, 2)); wh = r.uniform(0.1, 0.5, (B, n, 2))
box = torch.from_numpy(np.concatenate([c - wh / 2, c + wh / 2], 2).astype(np.float32)).cuda()
for fmt in ("f32", "u8", "bits"):
    out = None
    for _ in range(3): out, _ = assemble_masks_batch(proto, coef, box, size, size, True, fmt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): assemble_masks_batch(proto, coef, box, size, size, True, fmt, masks_out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("mask_assembly %s size %d B=%d n=100: %.4f ms  (%.0f GB/s of output)" % (fmt, size, B, ms, out.numel() * out.element_size() / ms / 1e6))
