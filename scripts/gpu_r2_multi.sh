#!/bin/bash
# Round 2, multi-GPU evidence (gpurun --gpus 8): the sharded-run == single-GPU equivalence test, then bench.py at N = 8 on
# BASELINE configs 2, 4 (im700, global batch 32) and 5 (plus_base, global batch 64).
mkdir -p gpurun_out; S=gpurun_out/r2multi_summary.txt; rm -f $S
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 >> $S
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2multi_test.log 2>&1; echo "test_gpu_multi exit $?" >> $S; tail -1 gpurun_out/r2multi_test.log >> $S
run() { tag=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $n --steps 10 --warmup 3 --no-fast-mode "$@" > gpurun_out/r2multi_$tag.log 2> gpurun_out/r2multi_$tag.err
  echo "bench $tag (N=$n) exit $?" >> $S
  python - gpurun_out/r2multi_$tag.log >> $S <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"metric"')][-1]
    print("  value %.0f FPS (%.3f ms/step, global batch %s)  e2e %.0f  e2e_bits %.0f" % (j["value"], j["ms_per_step"], j["config"]["global_batch"], j["e2e"]["value"], j["e2e_bits"]["value"]))
except Exception as e:
    print("  parse error", e)
PY
}
run base_n8 8
run im700_n8 8 --config yolact_im700_config --batch 4
run plus_base_n8 8 --config yolact_plus_base_config --batch 8
run base_n2 2
cat $S
