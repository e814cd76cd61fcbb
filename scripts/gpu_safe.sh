#!/bin/bash
# every step under a short timeout; a hung kernel traps after ~2 s (mbarrier watchdog)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in ${TESTS:-test_gpu_conv test_gpu_detect_post test_gpu_dcn test_gpu_network}; do
  timeout ${TEST_TIMEOUT:-300} python -m pytest tests/$t.py -m gpu -q -s -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  grep -E "^FAILED|passed|failed|timed out" gpurun_out/$t.log | cut -c1-200 | head -12 >> gpurun_out/summary.txt
done
if [ -z "$NO_LAYERS" ]; then
timeout 200 python scripts/layer_profile.py > gpurun_out/layers.md 2> gpurun_out/layers.err; echo "layers exit $?" >> gpurun_out/summary.txt
head -1 gpurun_out/layers.md >> gpurun_out/summary.txt; fi
if [ -z "$NO_BENCH" ]; then
timeout 400 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-2300 >> gpurun_out/summary.txt
tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt; fi
if [ -n "$EXTRA" ]; then bash -c "$EXTRA" >> gpurun_out/summary.txt 2>&1; fi
cat gpurun_out/summary.txt
