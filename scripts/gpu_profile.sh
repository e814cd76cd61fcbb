#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one steady-state step with its device time,
# (2) one --set full capture of the dominant kernels.  Never a bench number.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$R.csv python scripts/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:tc_conv -c ${NCU_COUNT:-12} -s ${NCU_SKIP:-20} -o gpurun_out/prof_tc_$R -f python scripts/profile_step.py --conv-only > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:mask_assembly -c 2 -o gpurun_out/prof_mask_$R -f python scripts/profile_step.py > gpurun_out/ncu_mask.log 2>&1
echo "mask capture exit $?"
ls -la gpurun_out | tail -20
