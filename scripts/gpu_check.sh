#!/bin/bash
# Runs on the GPU box via gpurun: per-file test logs + smoke + a short bench, all under their own timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
for t in test_gpu_conv test_gpu_detect_post test_gpu_dcn test_gpu_network; do
  timeout ${TEST_TIMEOUT:-600} python -m pytest tests/$t.py -m gpu -q -s -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$t.log >> gpurun_out/summary.txt
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
