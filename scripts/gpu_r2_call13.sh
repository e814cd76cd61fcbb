#!/bin/bash
# Round 2, call 13: generalised chain kernel (whole ResNet trunk in one launch: strides, 64-wide N tiles, DAG dependencies).
mkdir -p gpurun_out; S=gpurun_out/r2c13_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() { tag=$1; to=$2; shift 2; timeout $to "$@" > gpurun_out/r2c13_$tag.log 2>&1; echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c13_$tag.log | cut -c1-200 >> $S; grep -E "^FAILED|^ERROR|chains:" gpurun_out/r2c13_$tag.log | head -12 | cut -c1-300 >> $S; }
YB_LIB=$WD YB_CHAIN_VERBOSE=1 run wd_chain 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -s -p no:cacheprovider
grep -h "layers): chain" gpurun_out/r2c13_wd_chain.log | head -40 | cut -c1-200 >> $S
if grep -q "exit 124" $S; then cat $S; exit 0; fi
YB_LIB=$WD run wd_net 900 python -m pytest tests/test_gpu_network.py -m gpu -q -k "f16x3" -p no:cacheprovider
bench() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode "$@" > gpurun_out/r2c13_bench_$tag.log 2> gpurun_out/r2c13_bench_$tag.err
  echo "bench $tag exit $?" >> $S
  grep -h "layers): chain" gpurun_out/r2c13_bench_$tag.err | head -12 | cut -c1-200 >> $S
  python - "gpurun_out/r2c13_bench_$tag.log" >> $S <<'PY'
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
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3_chain.md 2>/dev/null; head -1 gpurun_out/layers_r02_f16x3_chain.md >> $S
run full 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x
# evidence: launch list of one steady-state step + one full capture of the chain kernel
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active"
YB_PRECISION=f16x3 timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/launches_r02_f16x3_chain.csv python scripts/profile_step.py > gpurun_out/ncu_launches_chain.log 2>&1
echo "launch list exit $?" >> $S
YB_PRECISION=f16x3 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"tc_chain" -c 3 \
    -o gpurun_out/prof_tc_chain_r02 -f python scripts/profile_step.py --conv-only > gpurun_out/ncu_tc_chain.log 2>&1
echo "ncu chain exit $?" >> $S
ncu -i gpurun_out/prof_tc_chain_r02.ncu-rep --page raw --csv > gpurun_out/prof_tc_chain_r02_raw.csv 2>/dev/null
ls -la gpurun_out/prof_tc_chain_r02.ncu-rep >> $S
[ $(stat -c %s gpurun_out/prof_tc_chain_r02.ncu-rep 2>/dev/null || echo 0) -gt 20000000 ] && rm -f gpurun_out/prof_tc_chain_r02.ncu-rep
cat $S
