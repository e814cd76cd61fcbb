#!/bin/bash
# Round 2, call 17: chain epilogue with the residual requested up front (registers) -- parity, wait statistics, bench, suite.
mkdir -p gpurun_out; S=gpurun_out/r2c17_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
YB_LIB=$WD timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c17_wd_chain.log 2>&1; echo "wd_chain exit $?" >> $S; tail -1 gpurun_out/r2c17_wd_chain.log >> $S
if grep -q "exit 124" $S; then cat $S; exit 0; fi
YB_CHAIN_STATS=1 YB_CHAIN_VERBOSE=1 timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/r2c17_layers_b8.md 2> gpurun_out/r2c17_b8.err
grep -h "chain\|groups" gpurun_out/r2c17_b8.err | cut -c1-700 >> $S
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode > gpurun_out/r2c17_bench.log 2> gpurun_out/r2c17_bench.err; echo "bench exit $?" >> $S
python - >> $S <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2c17_bench.log").read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f launches %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], j.get("gpu_launches")))
except Exception as e:
    print("  parse error", e)
PY
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2c17_full.log 2>&1; echo "full exit $?" >> $S; tail -1 gpurun_out/r2c17_full.log >> $S
cat $S
