#!/bin/bash
# Round 2 evidence run (1 GPU): ncu launch list with tensor-pipe / DRAM / L2 metrics for EVERY launch of one steady-state
# step (both precision modes), then one `--set full` capture per kernel of the path.  Summaries are written here by
# scripts/ncu_summary.py / summarize_launches.py into profiles/.
mkdir -p gpurun_out; S=gpurun_out/r2prof_summary.txt; rm -f $S
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active"
for prec in f16x3 f16tc; do
  YB_PRECISION=$prec timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
      --log-file gpurun_out/launches_r02_$prec.csv python scripts/profile_step.py > gpurun_out/ncu_launches_$prec.log 2>&1
  echo "launch list $prec exit $?" >> $S
done
YB_PRECISION=f16x3 timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/launches_r02_plus_f16x3.csv python scripts/profile_step.py --config yolact_plus_base_config > gpurun_out/ncu_launches_plus.log 2>&1
echo "launch list plus_base f16x3 exit $?" >> $S
# gpurun merges at most 64 MiB back: every report is exported to its raw-metrics CSV on the box (read here by
# scripts/ncu_csv_summary.py); only the small reports travel as .ncu-rep, the conv ones (40 MB each) are deleted.
cap() {  # cap <tag> <kernel regex> <count> <skip> <keep report: 0|1> <command...>
  tag=$1; re=$2; cnt=$3; skip=$4; keep=$5; shift 5
  timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"$re" -s $skip -c $cnt \
      -o gpurun_out/prof_${tag}_r02 -f "$@" > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag exit $?" >> $S
  ncu -i gpurun_out/prof_${tag}_r02.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_r02_raw.csv 2>/dev/null
  [ $keep = 1 ] || rm -f gpurun_out/prof_${tag}_r02.ncu-rep
}
export YB_PRECISION=f16x3
cap tc_split "tc_conv" 12 40 0 python scripts/profile_step.py --conv-only
cap stem "stem_tc" 1 0 1 python scripts/profile_step.py --conv-only
cap pointwise "maxpool|upsample" 4 0 0 python scripts/profile_step.py --conv-only
cap mask "mask_assembly" 1 0 1 python scripts/profile_step.py
cap detect "detect_candidates|class_nms|final_select" 3 0 1 python scripts/profile_step.py
cap dcn "dcn_tc" 3 0 1 python scripts/profile_step.py --config yolact_plus_base_config --conv-only
YB_PRECISION=f16tc cap tc_f16 "tc_conv" 12 40 0 python scripts/profile_step.py --conv-only
unset YB_PRECISION
timeout 600 ncu --set full --clock-control none -k regex:"trad_nms|mask_iou_bits|mask_rle|display_blend|fast_base_transform|pack_mask_bits|box_iou|pack_detections" -c 12 \
    -o gpurun_out/prof_rows_r02 -f python scripts/bench_rows.py > gpurun_out/ncu_rows.log 2>&1
echo "ncu rows exit $?" >> $S
ncu -i gpurun_out/prof_rows_r02.ncu-rep --page raw --csv > gpurun_out/prof_rows_r02_raw.csv 2>/dev/null
rm -f gpurun_out/prof_rows_r02.ncu-rep
ls -la gpurun_out/ >> $S 2>&1
du -sh gpurun_out >> $S
cat $S
