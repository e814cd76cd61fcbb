#!/bin/bash
# Round 2, last call: sanity of the committed tree (smoke, chain + network tests) and a fresh `--set full` capture of the
# chain kernel as built now (three stages).
mkdir -p gpurun_out; S=gpurun_out/r2last_summary.txt; rm -f $S
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('build+smoke ok')" > gpurun_out/r2last_smoke.log 2>&1; echo "build+smoke exit $?" >> $S
timeout 300 python -m pytest tests/test_gpu_chain.py tests/test_gpu_network.py -m gpu -q -p no:cacheprovider -k "f16x3" > gpurun_out/r2last_tests.log 2>&1; echo "chain+network tests exit $?" >> $S; tail -1 gpurun_out/r2last_tests.log >> $S
YB_PRECISION=f16x3 timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:"tc_chain" -c 2 \
    -o gpurun_out/prof_tc_chain3_r02 -f python scripts/profile_step.py --conv-only > gpurun_out/ncu_tc_chain3.log 2>&1
echo "ncu chain exit $?" >> $S
ncu -i gpurun_out/prof_tc_chain3_r02.ncu-rep --page raw --csv > gpurun_out/prof_tc_chain3_r02_raw.csv 2>/dev/null
rm -f gpurun_out/prof_tc_chain3_r02.ncu-rep
cat $S
