#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/call4_summary.txt; rm -f $S
timeout 200 python -m pytest tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -x -p no:cacheprovider > gpurun_out/post_tests.log 2>&1
echo "detect/post/eval-row tests exit $?" >> $S; tail -1 gpurun_out/post_tests.log >> $S
for tag in "BENCH_OVERLAP=0" "BENCH_OVERLAP=1" "BENCH_OVERLAP=1 YB_PDL=1"; do
  f=gpurun_out/bench_$(echo $tag | tr ' =' '__').log
  env $tag timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $f 2> $f.err; echo "bench [$tag] exit $?" >> $S
  python - "$f" >> $S <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f (%.3f ms)  conv %.3f ms %.0f TFLOP/s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["achieved"]))
except Exception as e:
    print("  parse error", e)
PY
  tail -2 $f.err >> $S
done
timeout 120 python - >> $S 2>&1 <<'PY'
import torch
x = torch.empty(968_000_000 // 4, dtype=torch.float32, device="cuda")
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("memset 968 MB (torch zero_): %.3f ms = %.0f GB/s" % (ms, 968e6 / ms / 1e6))
PY
timeout 120 python scripts/bench_rows.py > gpurun_out/rows2.md 2> gpurun_out/rows2.err; echo "rows exit $?" >> $S; grep -E "fast_base|display|rle" gpurun_out/rows2.md >> $S
cat $S
