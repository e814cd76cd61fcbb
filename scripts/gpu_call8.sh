#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/call8_summary.txt; rm -f $S
timeout 100 python scripts/diag_mask.py 550 8 >> $S 2>&1
timeout 100 python scripts/diag_mask.py 700 4 >> $S 2>&1
YB_MASK_CTAS_PER_SM=12 timeout 100 python scripts/diag_mask.py 700 4 >> $S 2>&1
timeout 100 python scripts/diag_mask.py 700 1 >> $S 2>&1
timeout 100 python scripts/diag_mask.py 640 4 >> $S 2>&1
BENCH_OVERLAP=0 timeout 200 python bench.py --steps 10 --warmup 3 --config yolact_im700_config --batch 4 --no-cpu-baseline > gpurun_out/im700_noov.log 2>&1
python - >> $S <<'PY'
import json
j=json.loads(open("gpurun_out/im700_noov.log").read().strip().splitlines()[-1])
print("im700 no-overlap: value %.0f (%.3f ms) e2e %.0f conv %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"]))
PY
cat $S
