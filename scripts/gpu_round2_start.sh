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
