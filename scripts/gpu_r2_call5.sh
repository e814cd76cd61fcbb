#!/bin/bash
# Round 2, call 5: Darknet padded layers + measured defaults -> whole gpu suite, smoke, bench (both arms), then the ncu
# evidence run (scripts/gpu_r2_profiles.sh).
mkdir -p gpurun_out; S=gpurun_out/r2c5_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() { tag=$1; to=$2; shift 2; timeout $to "$@" > gpurun_out/r2c5_$tag.log 2>&1; echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c5_$tag.log | cut -c1-200 >> $S; grep -E "^FAILED|^ERROR" gpurun_out/r2c5_$tag.log | head -8 | cut -c1-220 >> $S; }
YB_LIB=$WD run wd_net 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_eval_sequences.py -m gpu -q -p no:cacheprovider
if grep -q "exit 124" $S; then cat $S; exit 0; fi
run all_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c5_bench.log 2> gpurun_out/r2c5_bench.err; echo "bench exit $?" >> $S
tail -1 gpurun_out/r2c5_bench.log | cut -c1-3000 >> $S
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c5_bench_ref.log 2> gpurun_out/r2c5_bench_ref.err; echo "bench reference exit $?" >> $S
tail -1 gpurun_out/r2c5_bench_ref.log | cut -c1-600 >> $S
timeout 400 python scripts/parity_report.py --precisions f16x3,f16tc --out gpurun_out/parity_r02.json > gpurun_out/parity_r02.md 2> gpurun_out/parity_r02.err; echo "parity report exit $?" >> $S
for c in yolact_resnet50_config yolact_plus_resnet50_config yolact_im700_config yolact_plus_base_config yolact_darknet53_config; do
  b=8; [ $c = yolact_im700_config ] && b=4
  timeout 400 python bench.py --steps 10 --warmup 3 --config $c --batch $b --no-cpu-baseline > gpurun_out/r2c5_bench_$c.log 2> gpurun_out/r2c5_bench_$c.err
  echo "bench $c exit $?" >> $S
  python - gpurun_out/r2c5_bench_$c.log >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    f = j.get("fast_mode_f16tc") or {}
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f | f16tc value %.0f conv %.3f ms frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], f.get("value", 0), f.get("ms_conv_stack_per_step", 0), f.get("conv_frac_of_peak", 0)))
except Exception as e:
    print("  parse error", e)
PY
done
cat $S
bash scripts/gpu_r2_profiles.sh
