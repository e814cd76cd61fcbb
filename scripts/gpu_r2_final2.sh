#!/bin/bash
# Round 2, last validation of the committed tree: whole GPU suite, smoke, one bench line.
mkdir -p gpurun_out; S=gpurun_out/r2final2_summary.txt; rm -f $S
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2final2_full.log 2>&1; echo "pytest -m gpu exit $?" >> $S; tail -1 gpurun_out/r2final2_full.log >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2final2_full.log | head -8 | cut -c1-250 >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2final2_smoke.log 2>&1; echo "smoke exit $?" >> $S
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fast-mode > gpurun_out/r2final2_bench.json 2> gpurun_out/r2final2_bench.err; echo "bench exit $?" >> $S
python - >> $S <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2final2_bench.json").read().strip().splitlines()[-1])
    print("  value %.0f FPS (%.3f ms)  e2e %.0f  conv %.3f ms  frac %.3f launches %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["ms_conv_stack_per_step"], j["roofline"]["frac"], j.get("gpu_launches")))
except Exception as e:
    print("  parse error", e)
PY
cat $S
