#!/bin/bash
# Round 2, call 14: chain kernel with three split stages (32-channel staging tiles).
mkdir -p gpurun_out; S=gpurun_out/r2c14_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() { tag=$1; to=$2; shift 2; timeout $to "$@" > gpurun_out/r2c14_$tag.log 2>&1; echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c14_$tag.log | cut -c1-200 >> $S; grep -E "^FAILED|^ERROR|chains:" gpurun_out/r2c14_$tag.log | head -12 | cut -c1-300 >> $S; }
YB_LIB=$WD YB_CHAIN_VERBOSE=1 run wd_chain 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -s -p no:cacheprovider
grep -h "layers): chain" gpurun_out/r2c14_wd_chain.log | head -40 | cut -c1-200 >> $S
if grep -q "exit 124" $S; then cat $S; exit 0; fi
YB_LIB=$WD run wd_net 900 python -m pytest tests/test_gpu_network.py -m gpu -q -k "f16x3" -p no:cacheprovider
bench() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode "$@" > gpurun_out/r2c14_bench_$tag.log 2> gpurun_out/r2c14_bench_$tag.err
  echo "bench $tag exit $?" >> $S
  grep -h "layers): chain" gpurun_out/r2c14_bench_$tag.err | head -12 | cut -c1-200 >> $S
  python - "gpurun_out/r2c14_bench_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f launches %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], j.get("gpu_launches")))
except Exception as e:
    print("  parse error", e)
PY
}
bench chain_auto YB_CHAIN_VERBOSE=1 -- --precision f16x3
bench chain_forced YB_CHAIN=2 -- --precision f16x3
bench chain_off YB_CHAIN=0 -- --precision f16x3
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3_chain.md 2>/dev/null; head -1 gpurun_out/layers_r02_f16x3_chain.md >> $S
run full 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x
cat $S
