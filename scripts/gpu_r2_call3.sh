#!/bin/bash
# Round 2, call 3: bf16 lo planes (mixed bf16 x fp16 tcgen05 pass) -- diagnostic, tests, parity at full size, PDL A/B.
mkdir -p gpurun_out; S=gpurun_out/r2c3_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() {  # run <tag> <timeout> <cmd...>
  tag=$1; to=$2; shift 2
  timeout $to "$@" > gpurun_out/r2c3_$tag.log 2>&1
  echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c3_$tag.log | cut -c1-200 >> $S
  grep -E "^FAILED|^ERROR" gpurun_out/r2c3_$tag.log | head -8 | cut -c1-220 >> $S
}
YB_LIB=$WD run diag 120 python scripts/diag_split.py
cat gpurun_out/r2c3_diag.log >> $S
if grep -q "diag exit 124" $S; then cat $S; exit 0; fi
YB_LIB=$WD run wd_split_conv 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "split" -p no:cacheprovider
YB_LIB=$WD YB_TEST_EXPERIMENTAL=1 run wd_pdl 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "pdl" -p no:cacheprovider
YB_LIB=$WD run wd_dcn 300 python -m pytest tests/test_gpu_dcn.py -m gpu -q -s -p no:cacheprovider
YB_LIB=$WD run wd_net_plus 400 python -m pytest tests/test_gpu_network.py -m gpu -q -s -k "plus and (f16x3 or f16tc)" -p no:cacheprovider
if grep -q "exit 124" $S; then cat $S; exit 0; fi
run net_all 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_dcn.py tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -s -p no:cacheprovider
grep -E "f16x3 (proto|loc|conf|mask) rel err|f16x3 vs CPU" gpurun_out/r2c3_net_all.log | cut -c1-120 >> $S
run fullsize 900 python -m pytest tests/test_gpu_parity_fullsize.py tests/test_gpu_fullsize_goldens.py -m gpu -q -s -p no:cacheprovider
grep -E "raw_loc|strict class order|raw errors" gpurun_out/r2c3_fullsize.log | cut -c1-700 >> $S
bench() {  # bench <tag> <env...> -- <args...>
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode "$@" > gpurun_out/r2c3_bench_$tag.log 2> gpurun_out/r2c3_bench_$tag.err
  echo "bench $tag exit $?" >> $S
  python - "gpurun_out/r2c3_bench_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"]))
except Exception as e:
    print("  parse error", e)
PY
}
bench f16x3 A=1 -- --precision f16x3
bench f16tc A=1 -- --precision f16tc
if grep -q "wd_pdl exit 0" $S; then
  bench f16tc_pdl YB_PDL=1 -- --precision f16tc
  YB_PDL=1 timeout 300 python scripts/layer_profile.py --precision f16tc > gpurun_out/layers_r02_f16tc_pdl.md 2>/dev/null; echo "pdlf layers: $(grep -c pdlf gpurun_out/layers_r02_f16tc_pdl.md)" >> $S
fi
bench plus_f16x3_fused YB_DCN_FUSED=1 -- --precision f16x3 --config yolact_plus_base_config
bench plus_f16x3_unfused YB_DCN_FUSED=0 -- --precision f16x3 --config yolact_plus_base_config
bench plus_f16tc_fused YB_DCN_FUSED=1 -- --precision f16tc --config yolact_plus_base_config
bench plus_f16tc_unfused YB_DCN_FUSED=0 -- --precision f16tc --config yolact_plus_base_config
timeout 300 python scripts/layer_profile.py --precision f16tc --config yolact_plus_base_config > gpurun_out/layers_r02_plus_f16tc.md 2>/dev/null
timeout 300 python scripts/layer_profile.py --precision f16x3 --config yolact_plus_base_config > gpurun_out/layers_r02_plus_f16x3.md 2>/dev/null
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3.md 2>/dev/null; head -1 gpurun_out/layers_r02_f16x3.md >> $S
cat $S
