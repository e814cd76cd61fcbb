#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/call5_summary.txt; rm -f $S
timeout 200 python -m pytest tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -x -p no:cacheprovider > gpurun_out/post_tests.log 2>&1
echo "detect/post/eval-row tests exit $?" >> $S; tail -1 gpurun_out/post_tests.log >> $S; grep -E "^FAILED|Error" gpurun_out/post_tests.log | head -5 | cut -c1-250 >> $S
timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -x -p no:cacheprovider > gpurun_out/net_tests.log 2>&1
echo "network tests exit $?" >> $S; tail -1 gpurun_out/net_tests.log >> $S
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> $S
python - gpurun_out/bench.log >> $S <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value %.0f FPS (%.3f ms)  e2e %.0f (%.3f ms)  conv %.3f ms %.0f TFLOP/s frac %.3f cpu %s launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["achieved"], j["roofline"]["frac"], j.get("cpu_baseline"), j["gpu_launches"]))
PY
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01h.csv python scripts/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> $S
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:mask_assembly -c 1 -o gpurun_out/prof_mask_r01h -f python scripts/profile_step.py > gpurun_out/ncu_mask.log 2>&1
echo "mask capture exit $?" >> $S
cat $S
