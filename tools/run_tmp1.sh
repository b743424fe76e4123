o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x > $o/s25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/s25_pytest.log
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))
except Exception as e: print('$tag', 'ERR', e)
"; tail -3 $o/$tag.err; }
b s25_tail2 A=1
b s25_tail1 UB_TAIL_NODE=1
b s25_tail3 UB_TAIL_NODE=3
