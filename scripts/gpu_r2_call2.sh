#!/bin/bash
# Round 2, call 2: bring-up of the split-precision (f16x3) mode.  Watchdog build first (a protocol bug traps instead of
# hanging the GPU), then the release build: op tests, network tests, full-size parity, per-layer profile, bench A/B.
mkdir -p gpurun_out; S=gpurun_out/r2c2_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() {  # run <tag> <timeout> <cmd...>
  tag=$1; to=$2; shift 2
  timeout $to "$@" > gpurun_out/r2c2_$tag.log 2>&1
  echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c2_$tag.log | cut -c1-200 >> $S
  grep -E "^FAILED|^ERROR" gpurun_out/r2c2_$tag.log | head -6 | cut -c1-220 >> $S
}
YB_LIB=$WD run wd_split_conv 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "split" -p no:cacheprovider
if grep -q "wd_split_conv exit 124" $S; then cat $S; tail -30 gpurun_out/r2c2_wd_split_conv.log; exit 0; fi
YB_LIB=$WD run wd_dcn 200 python -m pytest tests/test_gpu_dcn.py -m gpu -q -x -p no:cacheprovider
YB_LIB=$WD run wd_net_small 600 python -m pytest tests/test_gpu_network.py -m gpu -q -k "exact_modes and f16x3" -s -p no:cacheprovider
YB_LIB=$WD YB_TEST_EXPERIMENTAL=1 run wd_pdl 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "pdl" -p no:cacheprovider
if grep -q "wd_net_small exit 124" $S; then cat $S; tail -40 gpurun_out/r2c2_wd_net_small.log; exit 0; fi
run conv_all 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider
run net_all 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_dcn.py tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -s -p no:cacheprovider
run fullsize 900 python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -q -s -p no:cacheprovider
grep -E "raw_loc" gpurun_out/r2c2_fullsize.log | cut -c1-600 >> $S
for prec in f16x3 f16tc; do
  timeout 300 python scripts/layer_profile.py --precision $prec > gpurun_out/layers_r02_$prec.md 2> gpurun_out/layers_r02_$prec.err; echo "layers $prec exit $?" >> $S
  head -1 gpurun_out/layers_r02_$prec.md >> $S
  timeout 400 python bench.py --steps 20 --warmup 5 --precision $prec --no-cpu-baseline --no-fast-mode > gpurun_out/r2c2_bench_$prec.log 2> gpurun_out/r2c2_bench_$prec.err
  echo "bench $prec exit $?" >> $S
  python - "gpurun_out/r2c2_bench_$prec.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"]))
except Exception as e:
    print("  parse error", e)
PY
done
if grep -q "wd_pdl exit 0" $S; then
  YB_PDL=1 timeout 400 python bench.py --steps 20 --warmup 5 --precision f16tc --no-cpu-baseline --no-fast-mode > gpurun_out/r2c2_bench_pdl.log 2> gpurun_out/r2c2_bench_pdl.err
  echo "bench f16tc YB_PDL=1 exit $?" >> $S
  python - gpurun_out/r2c2_bench_pdl.log >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"]))
except Exception as e:
    print("  parse error", e)
PY
fi
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c2_bench_ref.log 2> gpurun_out/r2c2_bench_ref.err; echo "bench reference exit $?" >> $S
tail -1 gpurun_out/r2c2_bench_ref.log | cut -c1-400 >> $S
tail -3 gpurun_out/r2c2_bench_f16x3.err >> $S
cat $S
