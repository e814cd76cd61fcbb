#!/bin/bash
# round-end style check: tests (short timeouts), smoke, bench, other configs, row kernels, DRAM bytes of the conv launches
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
for t in ${TESTS:-test_gpu_eval_rows test_gpu_detect_post test_gpu_conv test_gpu_dcn test_gpu_network}; do
  timeout ${TEST_TIMEOUT:-300} python -m pytest tests/$t.py -m gpu -q -x -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> $S
  grep -E "^FAILED|^ERROR|passed|failed|timed out|Error" gpurun_out/$t.log | cut -c1-300 | head -8 >> $S
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> $S; tail -2 gpurun_out/smoke.log >> $S
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> $S
tail -1 gpurun_out/bench.log | cut -c1-2500 >> $S; tail -3 gpurun_out/bench.err >> $S
timeout 120 python scripts/bench_rows.py > gpurun_out/rows.md 2> gpurun_out/rows.err; echo "rows exit $?" >> $S; cat gpurun_out/rows.md >> $S; tail -3 gpurun_out/rows.err >> $S
if [ -z "$NO_CONFIGS" ]; then
for c in yolact_resnet50_config yolact_plus_resnet50_config yolact_plus_base_config yolact_darknet53_config; do
  timeout 300 python bench.py --steps 10 --warmup 3 --config $c --no-cpu-baseline > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "bench $c exit $?" >> $S
  tail -1 gpurun_out/bench_$c.log | cut -c1-700 >> $S; tail -2 gpurun_out/bench_$c.err >> $S
done
timeout 300 python bench.py --steps 10 --warmup 3 --config yolact_im700_config --batch 4 --no-cpu-baseline > gpurun_out/bench_im700.log 2> gpurun_out/bench_im700.err; echo "bench im700 exit $?" >> $S
tail -1 gpurun_out/bench_im700.log | cut -c1-700 >> $S; tail -2 gpurun_out/bench_im700.err >> $S
fi
if [ -z "$NO_NCU" ]; then
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${ROUND:-r01}.csv python scripts/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> $S
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:tc_conv -c 12 -s 60 -o gpurun_out/prof_tc_${ROUND:-r01} -f python scripts/profile_step.py --conv-only > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?" >> $S
timeout 400 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/traffic_${ROUND:-r01}.csv python scripts/profile_step.py > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?" >> $S
fi
cat $S
