o=gpurun_out; mkdir -p $o
UB_MICROBATCH=2 UB_MICROBATCH_LOWRES=256 timeout 200 python -m pytest tests/test_trainer_gpu.py -q -x -k "B4 or B32 or eager_and_graph or ten_step" > $o/k1_pytest.log 2>&1; echo "pytest lowres rc=$?"; tail -2 $o/k1_pytest.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); print('$tag', 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'])
except Exception as e: print('$tag', 'ERR', e)
"; }
run k1_base UB_X=1
run k1_low256 UB_MICROBATCH=2 UB_MICROBATCH_LOWRES=256
run k1_low64 UB_MICROBATCH=2 UB_MICROBATCH_LOWRES=64
run k1_low256_fwd UB_MICROBATCH_FWD=2 UB_MICROBATCH_LOWRES=256
run k1_low64_fwd UB_MICROBATCH_FWD=2 UB_MICROBATCH_LOWRES=64
run k1_low1024 UB_MICROBATCH=2 UB_MICROBATCH_LOWRES=1024
run k1_base2 UB_X=1
