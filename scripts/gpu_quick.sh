#!/bin/bash
# quick check after a post-processing change: Detect / postprocess / row-kernel tests, network tests, bench, row timings
mkdir -p gpurun_out; S=gpurun_out/quick_summary.txt; rm -f $S
timeout 200 python -m pytest tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -x -p no:cacheprovider > gpurun_out/post_tests.log 2>&1
echo "detect/post/eval-row tests exit $?" >> $S; tail -1 gpurun_out/post_tests.log >> $S; grep -E "^FAILED|Error" gpurun_out/post_tests.log | head -5 | cut -c1-250 >> $S
timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -x -p no:cacheprovider > gpurun_out/net_tests.log 2>&1
echo "network tests exit $?" >> $S; tail -1 gpurun_out/net_tests.log >> $S
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench exit $?" >> $S
python - >> $S <<'PY'
import json
j=json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
print("  value %.0f FPS (%.3f ms)  e2e %.0f (%.3f ms)  conv %.3f ms %.0f TFLOP/s frac %.3f launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["achieved"], j["roofline"]["frac"], j["gpu_launches"]))
PY
timeout 120 python scripts/bench_rows.py > gpurun_out/rows_quick.md 2> gpurun_out/rows_quick.err; echo "rows exit $?" >> $S; grep -E "detect" gpurun_out/rows_quick.md >> $S
cat $S
