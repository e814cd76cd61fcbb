#!/bin/bash
# bring-up of the CTA-pair (cta_group::2) conv kernel: tests on the watchdog build first, then A/B numbers
mkdir -p gpurun_out; S=gpurun_out/pair_summary.txt; rm -f $S
WD=$PWD/yolact_b200/libyolact_b200_wd.so
YB_LIB=$WD timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "pair" -p no:cacheprovider > gpurun_out/pair_conv.log 2>&1
rc=$?; echo "pair conv tests (watchdog build) exit $rc" >> $S
grep -E "^FAILED|^ERROR|passed|failed|Error|error" gpurun_out/pair_conv.log | cut -c1-300 | head -12 >> $S
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pair_conv.log | cut -c1-250 >> $S; cat $S; exit 0; fi
timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "pair" -p no:cacheprovider > gpurun_out/pair_conv_rel.log 2>&1
echo "pair conv tests (release build) exit $?" >> $S; tail -1 gpurun_out/pair_conv_rel.log >> $S
YB_PAIR=1 timeout 300 python scripts/layer_profile.py > gpurun_out/layers_pair.md 2> gpurun_out/layers_pair.err; echo "layers pair exit $?" >> $S
head -1 gpurun_out/layers_pair.md >> $S; grep -c " pair " gpurun_out/layers_pair.md >> $S; tail -3 gpurun_out/layers_pair.err >> $S
YB_PAIR=1 timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pair_network.log 2>&1
echo "network tests with pair candidates exit $?" >> $S; tail -1 gpurun_out/pair_network.log >> $S
YB_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pair.log 2> gpurun_out/bench_pair.err; echo "bench pair exit $?" >> $S
tail -1 gpurun_out/bench_pair.log | cut -c1-1800 >> $S; tail -3 gpurun_out/bench_pair.err >> $S
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nopair.log 2> gpurun_out/bench_nopair.err; echo "bench nopair exit $?" >> $S
tail -1 gpurun_out/bench_nopair.log | cut -c1-600 >> $S
cat $S
