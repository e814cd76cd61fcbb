#!/bin/bash
# Round 2, call 19: chain producer reads the next tile's dependency counters ahead of time; 64-wide tiles for short layers.
mkdir -p gpurun_out; S=gpurun_out/r2c19_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
YB_LIB=$WD timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider -k "f16x3" > gpurun_out/r2c19_wd_chain.log 2>&1; echo "wd_chain exit $?" >> $S; tail -1 gpurun_out/r2c19_wd_chain.log >> $S
if grep -q "exit 124" $S; then cat $S; exit 0; fi
YB_LIB=$WD YB_CHAIN_BN64=1 timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider -k "f16x3 and (550 or 160)" > gpurun_out/r2c19_wd_chain64.log 2>&1; echo "wd_chain bn64 exit $?" >> $S; tail -1 gpurun_out/r2c19_wd_chain64.log >> $S
bench() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" YB_CHAIN_VERBOSE=1 YB_CHAIN_STATS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode "$@" > gpurun_out/r2c19_bench_$tag.log 2> gpurun_out/r2c19_bench_$tag.err
  echo "bench $tag exit $?" >> $S
  grep -h "layers): chain\|chain stats backbone" gpurun_out/r2c19_bench_$tag.err | cut -c1-420 >> $S
  python - "gpurun_out/r2c19_bench_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms  frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"]))
except Exception as e:
    print("  parse error", e)
PY
}
bench prepoll A=1 --
bench noprepoll YB_CHAIN_PREPOLL=0 --
bench bn64 YB_CHAIN_BN64=1 --
cat $S
