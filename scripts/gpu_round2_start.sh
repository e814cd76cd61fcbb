#!/bin/bash
# First GPU call of round 2: the data the round-1 budget did not reach.
#   * -m gpu tests added after the last round-1 GPU run (mask kernel at 8x / 16x resize, wide RLE)
#   * per-layer profiles of the configs whose conv stack is furthest from the base config (DCN, Darknet)
#   * ncu --set full captures, one launch each, of the kernels without one: DCN gather, mask assembly (final version),
#     traditional NMS, RLE, mask IoU, display blend
mkdir -p gpurun_out; S=gpurun_out/r2start_summary.txt; rm -f $S
timeout 300 python -m pytest tests/test_gpu_zz_upscale.py tests/test_gpu_detect_post.py tests/test_gpu_eval_rows.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1
echo "new + post tests exit $?" >> $S; tail -1 gpurun_out/r2_tests.log >> $S; grep -E "^FAILED|Error" gpurun_out/r2_tests.log | head -8 | cut -c1-250 >> $S
for c in yolact_plus_base_config yolact_plus_resnet50_config yolact_darknet53_config; do
  timeout 300 python scripts/layer_profile.py --config $c > gpurun_out/layers_$c.md 2> gpurun_out/layers_$c.err; echo "layers $c exit $?" >> $S
  head -1 gpurun_out/layers_$c.md >> $S
done
# experimental (written after the round-1 GPU budget was spent): PDL-friendly plans.  Tests on the watchdog build first.
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
# experimental: warp-cooperative DCN gather (YB_DCN_GATHER=warp)
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
cap() {  # cap <tag> <kernel regex> <command...>
  tag=$1; re=$2; shift 2
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"$re" -c 1 -o gpurun_out/prof_${tag}_r02 -f "$@" > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag exit $?" >> $S
}
cap dcn "dcn_gather" python scripts/profile_step.py --config yolact_plus_base_config --conv-only
cap mask "mask_assembly" python scripts/profile_step.py
cap tradnms "trad_nms" python scripts/bench_rows.py
cap rle "mask_rle" python scripts/bench_rows.py
cap maskiou "mask_iou_bits" python scripts/bench_rows.py
cap blend "display_blend" python scripts/bench_rows.py
cat $S
