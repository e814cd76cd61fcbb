#!/bin/bash
# Round 2, call 11: validation of the build after the A-stationary / epi3 variants were removed (res_direct kept).
mkdir -p gpurun_out; S=gpurun_out/r2c11_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
run() { tag=$1; to=$2; shift 2; timeout $to "$@" > gpurun_out/r2c11_$tag.log 2>&1; echo "$tag exit $?" >> $S; tail -1 gpurun_out/r2c11_$tag.log | cut -c1-200 >> $S; grep -E "^FAILED|^ERROR" gpurun_out/r2c11_$tag.log | head -8 | cut -c1-220 >> $S; }
YB_LIB=$WD run wd_conv 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider
if grep -q "exit 124" $S; then cat $S; exit 0; fi
run full 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; echo "bench exit $?" >> $S
python - >> $S <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2c11_bench.json").read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms  frac %.3f  fast %s  cpu %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], j.get("fast_mode_f16tc", {}).get("value"), j["cpu_baseline"]["value"]))
except Exception as e:
    print("  parse error", e)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c11_bench_ref.json 2> gpurun_out/r2c11_bench_ref.err; echo "bench ref exit $?" >> $S; tail -c 400 gpurun_out/r2c11_bench_ref.json >> $S
timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/layers_r02_f16x3.md 2>/dev/null; head -1 gpurun_out/layers_r02_f16x3.md >> $S
cat $S
