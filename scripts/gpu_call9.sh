#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/call9_summary.txt; rm -f $S
for cfg in "yolact_im700_config --batch 4" "yolact_base_config" "yolact_plus_base_config"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --config $cfg --no-cpu-baseline > gpurun_out/c9.log 2> gpurun_out/c9.err
  python - "$cfg" >> $S <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/c9.log").read().strip().splitlines()[-1])
    print("%s: value %.0f (%.3f ms) e2e %.0f (%.3f ms) conv %.3f" % (sys.argv[1], j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["ms_conv_stack_per_step"]))
except Exception as e:
    print(sys.argv[1], "error", e, open("gpurun_out/c9.err").read()[-400:])
PY
done
cat $S
