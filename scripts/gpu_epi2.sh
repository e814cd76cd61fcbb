#!/bin/bash
# bring-up of the two-epilogue-group conv variant and the 2-worker-group stem: tests first (watchdog build), then A/B
mkdir -p gpurun_out; S=gpurun_out/epi2_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
YB_LIB=$WD timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "epi2" -p no:cacheprovider > gpurun_out/epi2_conv.log 2>&1
rc=$?; echo "epi2 conv tests (watchdog build) exit $rc" >> $S
grep -E "^FAILED|^ERROR|passed|failed|Error|error" gpurun_out/epi2_conv.log | cut -c1-300 | head -12 >> $S
if [ $rc -ne 0 ]; then tail -40 gpurun_out/epi2_conv.log | cut -c1-250 >> $S; fi
timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider > gpurun_out/conv_all.log 2>&1
echo "all conv tests (release build) exit $?" >> $S; tail -1 gpurun_out/conv_all.log >> $S
YB_STEM_WG=2 timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -x -p no:cacheprovider > gpurun_out/net_stem2.log 2>&1
srcx=$?; echo "network tests, stem wg=2 exit $srcx" >> $S; tail -1 gpurun_out/net_stem2.log >> $S
E=""
if [ $rc -eq 0 ]; then
  YB_EPI2=1 timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -x -p no:cacheprovider > gpurun_out/net_epi2.log 2>&1
  echo "network tests, epi2 candidates exit $?" >> $S; tail -1 gpurun_out/net_epi2.log >> $S
  E="YB_EPI2=1"
fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err; echo "bench default exit $?" >> $S
tail -1 gpurun_out/bench_a.log | cut -c1-1700 >> $S
if [ $rc -eq 0 ]; then
env YB_EPI2=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err; echo "bench YB_EPI2=1 exit $?" >> $S
tail -1 gpurun_out/bench_b.log | cut -c1-1700 >> $S; tail -2 gpurun_out/bench_b.err >> $S
fi
if [ $srcx -eq 0 ]; then
env $E YB_STEM_WG=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c.log 2> gpurun_out/bench_c.err; echo "bench $E YB_STEM_WG=2 exit $?" >> $S
tail -1 gpurun_out/bench_c.log | cut -c1-1700 >> $S; tail -2 gpurun_out/bench_c.err >> $S
env $E YB_STEM_WG=2 timeout 300 python scripts/layer_profile.py > gpurun_out/layers_c.md 2> gpurun_out/layers_c.err; echo "layers exit $?" >> $S
head -4 gpurun_out/layers_c.md | tail -1 >> $S; grep -c "epi2" gpurun_out/layers_c.md >> $S
fi
cat $S
