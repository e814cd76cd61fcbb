#!/bin/bash
# Round 2, call 1: (a) full-size parity figures of the round-1 f16tc mode (the baseline the split-precision mode must beat),
# (b) validation of the two never-run experimental paths: PDL-friendly plans (watchdog build first) and the warp DCN gather.
mkdir -p gpurun_out; S=gpurun_out/r2c1_summary.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader >> $S
timeout 900 python scripts/parity_report.py --precisions f16tc --cases 0,1,2,3,4 --out gpurun_out/parity_f16tc_r02.json > gpurun_out/parity_f16tc_r02.md 2> gpurun_out/parity_f16tc_r02.err
echo "parity f16tc exit $?" >> $S; cat gpurun_out/parity_f16tc_r02.md >> $S
python -m yolact_b200.build --watchdog > /dev/null 2>&1
YB_LIB=$PWD/yolact_b200/libyolact_b200_wd.so YB_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "pdl" -p no:cacheprovider > gpurun_out/r2_pdl_tests.log 2>&1
rc=$?; echo "pdl-friendly conv tests (watchdog build) exit $rc" >> $S; tail -1 gpurun_out/r2_pdl_tests.log >> $S
if [ $rc -eq 0 ]; then
  for tag in "YB_PDL=0" "YB_PDL=1"; do
    env $tag timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_$tag.log 2> gpurun_out/r2_bench_$tag.err
    echo "bench [$tag] exit $?" >> $S
    python - "gpurun_out/r2_bench_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"]))
except Exception as e:
    print("  parse error", e)
PY
  done
  YB_PDL=1 timeout 300 python scripts/layer_profile.py > gpurun_out/layers_pdl.md 2> gpurun_out/layers_pdl.err; echo "layers pdl exit $?; pdlf layers: $(grep -c pdlf gpurun_out/layers_pdl.md)" >> $S
fi
YB_DCN_GATHER=warp timeout 300 python -m pytest tests/test_gpu_dcn.py tests/test_gpu_network.py -m gpu -q -k "dcn or plus" -p no:cacheprovider > gpurun_out/r2_dcn_warp_tests.log 2>&1
rc=$?; echo "dcn warp-gather tests exit $rc" >> $S; tail -1 gpurun_out/r2_dcn_warp_tests.log >> $S
for tag in "YB_DCN_GATHER=thread" "YB_DCN_GATHER=warp"; do
  env $tag timeout 300 python bench.py --steps 10 --warmup 3 --config yolact_plus_base_config --no-cpu-baseline > gpurun_out/r2_plus_$tag.log 2> gpurun_out/r2_plus_$tag.err
  echo "bench plus_base [$tag] exit $?" >> $S
  python - "gpurun_out/r2_plus_$tag.log" >> $S <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  conv %.3f ms" % (j["value"], j["ms_per_step"], j["roofline"]["ms_conv_stack_per_step"]))
except Exception as e:
    print("  parse error", e)
PY
done
YB_DCN_GATHER=warp timeout 200 python scripts/layer_profile.py --config yolact_plus_base_config > gpurun_out/layers_plus_warp.md 2>/dev/null
timeout 200 python scripts/layer_profile.py --config yolact_plus_base_config > gpurun_out/layers_plus_thread.md 2>/dev/null
cat $S
