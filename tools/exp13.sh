#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-reference-cuda --steps 40 --warmup 10"
try() { tag=$1; n=$2; shift; shift; for i in $(seq $n); do env "$@" timeout 50 $B > $o/x13_$tag$i.json 2> $o/x13_$tag$i.err; echo "$tag#$i rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$o/x13_$tag$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('loss_after'))
except Exception as e: print('ERR', e)
PY
)"; done; }
try default 2 UB_X=0
timeout 100 python tools/profile_ops.py > $o/x13_ops.txt 2>&1
awk -F'\t' '$3==5 || ($1=="bwd" && $2>=255) {print}' $o/ops.tsv | head -30
