#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in ${TESTS:-test_gpu_conv test_gpu_network}; do
  timeout ${TEST_TIMEOUT:-900} python -m pytest tests/$t.py -m gpu -q -s -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  grep -E "^FAILED|passed|failed" gpurun_out/$t.log | cut -c1-160 >> gpurun_out/summary.txt
done
timeout 600 python scripts/layer_profile.py > gpurun_out/layers.md 2> gpurun_out/layers.err; echo "layers exit $?" >> gpurun_out/summary.txt
head -3 gpurun_out/layers.md >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-2200 >> gpurun_out/summary.txt
tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
