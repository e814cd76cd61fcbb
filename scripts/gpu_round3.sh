#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in ${TESTS:-test_gpu_conv test_gpu_dcn test_gpu_network}; do
  timeout ${TEST_TIMEOUT:-900} python -m pytest tests/$t.py -m gpu -q -s -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  grep -E "^FAILED|passed|failed" gpurun_out/$t.log | cut -c1-160 >> gpurun_out/summary.txt
done
timeout 600 python scripts/layer_profile.py > gpurun_out/layers.md 2> gpurun_out/layers.err; echo "layers exit $?" >> gpurun_out/summary.txt
head -3 gpurun_out/layers.md >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-1800 >> gpurun_out/summary.txt
tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
YB_PDL=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl.log 2> gpurun_out/bench_pdl.err; echo "bench pdl exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_pdl.log | cut -c1-900 >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_pdl.err >> gpurun_out/summary.txt
YB_PDL=1 timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -p no:cacheprovider -k "f16tc" > gpurun_out/net_pdl.log 2>&1; echo "net pdl exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/net_pdl.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
