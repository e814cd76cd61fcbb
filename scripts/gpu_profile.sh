#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one steady-state step with its device time,
# (2) one --set full capture of the dominant kernels (cold cache, ncu default), (3) the mask kernel.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$R.csv python scripts/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:tc_conv -c ${NCU_COUNT:-14} -s ${NCU_SKIP:-18} -o gpurun_out/prof_tc_$R -f python scripts/profile_step.py --conv-only > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:mask_assembly -c 1 -o gpurun_out/prof_mask_$R -f python scripts/profile_step.py > gpurun_out/ncu_mask.log 2>&1
echo "mask capture exit $?"
timeout 400 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"class_nms|detect_candidates|final_select|stem_tc" -c 4 -o gpurun_out/prof_misc_$R -f python scripts/profile_step.py > gpurun_out/ncu_misc.log 2>&1
echo "misc capture exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$R.csv
