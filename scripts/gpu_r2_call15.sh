#!/bin/bash
# Round 2, call 15: where do the chain kernel's CTAs wait?  (YB_CHAIN_STATS: clock64 around every wait, per CTA)
mkdir -p gpurun_out; S=gpurun_out/r2c15_summary.txt; rm -f $S
YB_CHAIN_STATS=1 YB_CHAIN_VERBOSE=1 timeout 300 python scripts/layer_profile.py --precision f16x3 > gpurun_out/r2c15_layers_b8.md 2> gpurun_out/r2c15_b8.err
grep -h "chain" gpurun_out/r2c15_b8.err | cut -c1-700 >> $S
YB_CHAIN_STATS=1 YB_CHAIN_VERBOSE=1 timeout 300 python scripts/layer_profile.py --precision f16x3 --batch 2 > gpurun_out/r2c15_layers_b2.md 2> gpurun_out/r2c15_b2.err
grep -h "chain" gpurun_out/r2c15_b2.err | cut -c1-700 >> $S
cat $S
