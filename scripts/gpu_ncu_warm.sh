#!/bin/bash
# warm-cache ncu capture of stage-3 style kernels (cache-control none keeps L2 contents between replays)
mkdir -p gpurun_out
R=${ROUND:-r01c}
timeout 1500 ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on \
    -k regex:tc_conv -s ${NCU_SKIP:-30} -c ${NCU_COUNT:-9} -o gpurun_out/prof_tc_warm_$R -f python scripts/profile_step.py --conv-only > gpurun_out/ncu_warm.log 2>&1
echo "warm capture exit $?"
ls -la gpurun_out/*.ncu-rep
