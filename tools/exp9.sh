#!/bin/bash
o=gpurun_out; mkdir -p $o
T=unet.cu_b200/build/igemm_trace
timeout 200 unet.cu_b200/build/igemm_test > $o/x9_igemm_test.log 2>&1; echo "igemm_test rc=$? fails=$(grep -c FAIL $o/x9_igemm_test.log)"
for sh in "32 8 8 256 256" "32 16 16 192 192" "32 8 8 512 256"; do
  for n in 1 2; do echo "== $sh nacc<=$n +stats"; UB_TEST_STATS=1 UB_CONV_NACC=$n timeout 60 $T shape $sh 0 20; done
done > $o/x9_trace.txt 2>&1
grep -E "^==|last MMA|^conv" $o/x9_trace.txt
B="python bench.py --no-cpu-baseline --no-reference-cuda --steps 40 --warmup 10"
try() { tag=$1; n=$2; shift; shift; for i in $(seq $n); do env "$@" timeout 50 $B > $o/x9_$tag$i.json 2> $o/x9_$tag$i.err; echo "$tag#$i rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$o/x9_$tag$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('loss_after'))
except Exception as e: print('ERR', e)
PY
)"; done; }
try default 4 UB_X=0
try nacc1 2 UB_CONV_NACC=1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/profile_ops.py > $o/x9_ops.txt 2>&1
