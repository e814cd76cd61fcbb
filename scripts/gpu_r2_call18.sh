#!/bin/bash
# Round 2, call 18: chain kernel in the fp16 mode too (six operand stages); full default bench line.
mkdir -p gpurun_out; S=gpurun_out/r2c18_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
YB_LIB=$WD timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c18_wd_chain.log 2>&1; echo "wd_chain exit $?" >> $S; tail -1 gpurun_out/r2c18_wd_chain.log >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2c18_wd_chain.log | head -8 | cut -c1-250 >> $S
if grep -q "exit 124" $S; then cat $S; exit 0; fi
YB_LIB=$WD timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c18_wd_net.log 2>&1; echo "wd_net exit $?" >> $S; tail -1 gpurun_out/r2c18_wd_net.log >> $S
YB_CHAIN_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; echo "bench exit $?" >> $S
grep -h "layers): chain" gpurun_out/r2c18_bench.err | cut -c1-220 >> $S
python - >> $S <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2c18_bench.json").read().strip().splitlines()[-1])
    f = j.get("fast_mode_f16tc", {})
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms  frac %.3f  launches %s | fast: %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], j.get("gpu_launches"), {k: f.get(k) for k in ("value", "ms_per_step", "ms_conv_stack_per_step", "e2e")}))
except Exception as e:
    print("  parse error", e)
PY
YB_CHAIN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision f16tc --no-fast-mode > gpurun_out/r2c18_bench_f16tc_nochain.json 2>/dev/null
python - >> $S <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2c18_bench_f16tc_nochain.json").read().strip().splitlines()[-1])
    print("  f16tc YB_CHAIN=0: value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"]))
except Exception as e:
    print("  parse error", e)
PY
cat $S
