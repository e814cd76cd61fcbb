#!/bin/bash
f() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f e2e %.0f conv_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['ms_conv_stack_per_step']))"; }
echo "A base(babe48a):"; (cd alt/base && timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | f)
echo "B head branches=0 mrep=0:"; YB_BRANCHES=0 YB_MREP2=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | f
echo "C head nowd branches=0 mrep=0:"; YB_LIB=$PWD/alt/nowd/libyolact_b200.so YB_BRANCHES=0 YB_MREP2=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | f
echo "D head default:"; timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | f
echo "A2 base again:"; (cd alt/base && timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | f)
