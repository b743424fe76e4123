#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 200 unet.cu_b200/build/igemm_test quick > $o/x10_igemm_test.log 2>&1; echo "igemm_test rc=$? fails=$(grep -c FAIL $o/x10_igemm_test.log)"
B="python bench.py --no-cpu-baseline --no-reference-cuda --steps 40 --warmup 10"
try() { tag=$1; n=$2; shift; shift; for i in $(seq $n); do env "$@" timeout 50 $B > $o/x10_$tag$i.json 2> $o/x10_$tag$i.err; echo "$tag#$i rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$o/x10_$tag$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('loss_after'))
except Exception as e: print('ERR', e)
PY
)"; done; }
try default 3 UB_X=0
try nacc1 1 UB_CONV_NACC=1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/profile_ops.py > $o/x10_ops.txt 2>&1
