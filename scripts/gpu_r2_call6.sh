#!/bin/bash
# Round 2, call 6: stem worker-group A/B, the bench lines in full, the parity report, then the ncu evidence run.
mkdir -p gpurun_out; S=gpurun_out/r2c6_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
YB_LIB=$WD YB_STEM_WG=4 timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -k "raw_heads_exact_modes and f16x3" -p no:cacheprovider > gpurun_out/r2c6_wd_wg4.log 2>&1; echo "wd stem wg4 exit $?" >> $S; tail -1 gpurun_out/r2c6_wd_wg4.log >> $S
bench() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/r2c6_bench_$tag.log 2> gpurun_out/r2c6_bench_$tag.err
  echo "bench $tag exit $?" >> $S
  python - "gpurun_out/r2c6_bench_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    f = j.get("fast_mode_f16tc") or {}
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %.0f  conv %.3f ms  frac %.3f | f16tc value %.0f conv %.3f ms frac %.3f | cpu %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_bits"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], f.get("value", 0), f.get("ms_conv_stack_per_step", 0), f.get("conv_frac_of_peak", 0), (j.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("  parse error", e)
PY
}
bench wg2 A=1 -- --no-cpu-baseline --no-fast-mode
bench wg4 YB_STEM_WG=4 -- --no-cpu-baseline --no-fast-mode
bench wg1 YB_STEM_WG=1 -- --no-cpu-baseline --no-fast-mode
bench default A=1 --
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c6_bench_ref.log 2> gpurun_out/r2c6_bench_ref.err; echo "bench reference exit $?" >> $S
timeout 600 python scripts/parity_report.py --precisions f16x3,f16tc --out gpurun_out/parity_r02.json > gpurun_out/parity_r02.md 2> gpurun_out/parity_r02.err; echo "parity report exit $?" >> $S
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3.md 2>/dev/null
timeout 300 python scripts/layer_profile.py --precision f16tc > gpurun_out/layers_r02_f16tc.md 2>/dev/null
timeout 300 python scripts/bench_rows.py > gpurun_out/rows_r02.md 2>/dev/null
cat $S
bash scripts/gpu_r2_profiles.sh
