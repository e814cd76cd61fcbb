#!/bin/bash
# Round 2, 2-GPU call on the final (chain) build: sharded run == single-GPU run, bench.py at N = 2.
mkdir -p gpurun_out; S=gpurun_out/r2multi2_summary.txt; rm -f $S
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 >> $S
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2multi2_test.log 2>&1; echo "test_gpu_multi exit $?" >> $S; tail -1 gpurun_out/r2multi2_test.log >> $S
grep -E "^FAILED|^ERROR|Error" gpurun_out/r2multi2_test.log | head -6 | cut -c1-250 >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-fast-mode > gpurun_out/bench_r02_base_n2.json 2> gpurun_out/r2multi2_bench.err
echo "bench N=2 exit $?" >> $S
python - >> $S <<'PY'
import json
try:
    j = [json.loads(l) for l in open("gpurun_out/bench_r02_base_n2.json") if l.startswith('{"metric"')][-1]
    print("  value %.0f FPS (%.3f ms/step, global batch %s)  e2e %.0f  e2e_bits %.0f" % (j["value"], j["ms_per_step"], j["config"]["global_batch"], j["e2e"]["value"], j["e2e_bits"]["value"]))
except Exception as e:
    print("  parse error", e)
PY
cat $S
