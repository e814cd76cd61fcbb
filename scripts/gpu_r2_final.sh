#!/bin/bash
# Round 2, final 1-GPU evidence run on the final build: whole GPU suite, the bench lines (both arms, every BASELINE config),
# the full-size parity report, per-op times.
mkdir -p gpurun_out; S=gpurun_out/r2final_summary.txt; rm -f $S
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2final_full.log 2>&1; echo "pytest -m gpu exit $?" >> $S; tail -1 gpurun_out/r2final_full.log >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2final_full.log | head -8 | cut -c1-250 >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_base_n1.json 2> gpurun_out/r2final_bench.err; echo "bench exit $?" >> $S
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_n1.json 2> gpurun_out/r2final_ref.err; echo "bench reference exit $?" >> $S
line() { python - "$1" "$2" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    f = j.get("fast_mode_f16tc") or {}
    print("  %-28s value %.0f FPS (%.3f ms)  e2e %.0f  e2e_bits %s  conv %s ms  frac %s  launches %s  fast %s  cpu %s" % (
        sys.argv[2], j["value"], j["ms_per_step"], j["e2e"]["value"], (j.get("e2e_bits") or {}).get("value"),
        (j.get("roofline") or {}).get("ms_conv_stack_per_step"), (j.get("roofline") or {}).get("frac"), j.get("gpu_launches"),
        f.get("value"), (j.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("  parse error", sys.argv[2], e)
PY
}
line gpurun_out/bench_r02_base_n1.json yolact_base
line gpurun_out/bench_r02_reference_n1.json reference_arm
for spec in "yolact_resnet50_config 8" "yolact_plus_resnet50_config 8" "yolact_im700_config 4" "yolact_plus_base_config 8" "yolact_darknet53_config 8"; do
  set -- $spec
  timeout 600 python bench.py --config $1 --batch $2 --steps 10 --warmup 3 --no-cpu-baseline --no-fast-mode > gpurun_out/bench_r02_$1.json 2> gpurun_out/r2final_$1.err
  line gpurun_out/bench_r02_$1.json $1
done
timeout 900 python scripts/parity_report.py --precisions f16x3 --out gpurun_out/parity_r02_final.json > gpurun_out/parity_r02_final.md 2> gpurun_out/parity_r02_final.err; echo "parity report exit $?" >> $S
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3_final.md 2>/dev/null; head -1 gpurun_out/layers_r02_f16x3_final.md >> $S
cat $S
