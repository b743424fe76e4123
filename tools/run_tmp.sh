o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x > $o/s24_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/s24_pytest.log
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['profile_total_ms'],3), {k:(round(v['ms'],3), v.get('tflops')) for k,v in d['kernel_classes'].items()})
except Exception as e: print('$tag', 'ERR', e)
"; tail -3 $o/$tag.err; }
b s24_new A=1
b s24_wg128 UB_WGRAD_SMEM_KB=128
b s24_wg64 UB_WGRAD_SMEM_KB=64
timeout 120 python tools/timeline.py 32 $o/s24_timeline.tsv > $o/s24_timeline.txt 2>&1
