#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/call6_summary.txt; rm -f $S
timeout 200 python -m pytest tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -x -p no:cacheprovider > gpurun_out/post_tests.log 2>&1
echo "detect/post/eval-row tests exit $?" >> $S; tail -1 gpurun_out/post_tests.log >> $S; grep -E "^FAILED|Error" gpurun_out/post_tests.log | head -5 | cut -c1-250 >> $S
cat > /tmp/mask_time.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from yolact_b200.output_utils import assemble_masks_batch
r = np.random.RandomState(0)
B, n = 8, 100
proto = torch.from_numpy(np.maximum(r.standard_normal((B, 138, 138, 32)), 0).astype(np.float32)).cuda()
coef = torch.from_numpy(np.tanh(r.standard_normal((B, n, 32))).astype(np.float32)).cuda()
c = r.uniform(0.2, 0.8, (B, n, 2)); wh = r.uniform(0.1, 0.5, (B, n, 2))
box = torch.from_numpy(np.concatenate([c - wh / 2, c + wh / 2], 2).astype(np.float32)).cuda()
for fmt in ("f32", "u8", "bits"):
    out = None
    for _ in range(3): out, _ = assemble_masks_batch(proto, coef, box, 550, 550, True, fmt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): assemble_masks_batch(proto, coef, box, 550, 550, True, fmt, masks_out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("mask_assembly %s B=8 n=100 550^2: %.4f ms  (%.0f GB/s of output)" % (fmt, ms, out.numel() * out.element_size() / ms / 1e6))
PY
for c in 12 32 64; do echo "YB_MASK_CTAS_PER_SM=$c" >> $S; YB_MASK_CTAS_PER_SM=$c timeout 120 python /tmp/mask_time.py >> $S 2>&1; done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> $S
python - gpurun_out/bench.log >> $S <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value %.0f FPS (%.3f ms)  e2e %.0f (%.3f ms)  conv %.3f ms %.0f TFLOP/s frac %.3f launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["achieved"], j["roofline"]["frac"], j["gpu_launches"]))
PY
timeout 120 python scripts/bench_rows.py > gpurun_out/rows3.md 2> gpurun_out/rows3.err; echo "rows exit $?" >> $S; grep -E "rle" gpurun_out/rows3.md >> $S
cat $S
