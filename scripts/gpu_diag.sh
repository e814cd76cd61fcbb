#!/bin/bash
mkdir -p gpurun_out
timeout 100 env YB_AUTOTUNE=0 YB_BRANCHES=0 YB_TRACE=1 python -u scripts/diag.py > gpurun_out/diag.log 2>&1
tail -30 gpurun_out/diag.log | cut -c1-200
