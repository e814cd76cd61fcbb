#!/bin/bash
# Round 2, call 10: epilogue buffering variants (epi3, split epi2 with the residual read from global memory).
mkdir -p gpurun_out; S=gpurun_out/r2c10_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() { tag=$1; to=$2; shift 2; timeout $to "$@" > gpurun_out/r2c10_$tag.log 2>&1; echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c10_$tag.log | cut -c1-200 >> $S; grep -E "^FAILED|^ERROR" gpurun_out/r2c10_$tag.log | head -8 | cut -c1-220 >> $S; }
YB_LIB=$WD run wd_conv 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "split" -p no:cacheprovider
if grep -q "exit 124" $S; then cat $S; exit 0; fi
timeout 300 python scripts/probe_epilogue.py > gpurun_out/probe_epilogue2.txt 2>&1; cat gpurun_out/probe_epilogue2.txt >> $S
YB_LIB=$WD run wd_net 600 python -m pytest tests/test_gpu_network.py -m gpu -q -k "f16x3" -p no:cacheprovider
bench() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode "$@" > gpurun_out/r2c10_bench_$tag.log 2> gpurun_out/r2c10_bench_$tag.err
  echo "bench $tag exit $?" >> $S
  python - "gpurun_out/r2c10_bench_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"]))
except Exception as e:
    print("  parse error", e)
PY
}
bench f16x3 A=1 -- --precision f16x3
bench f16x3_noepi2 YB_EPI2=0 -- --precision f16x3
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3.md 2>/dev/null; head -1 gpurun_out/layers_r02_f16x3.md >> $S
cat $S
