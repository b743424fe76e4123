o=gpurun_out; mkdir -p $o
# n3: end-of-CTA bulk waits: .read only (default) vs full completion (UB_LIB_VARIANT=fullwait)
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); k=d['kernel_classes']; print('$tag', 'ms', round(d['ms_per_step'],4), 'conv', k['conv_igemm']['ms'], 'wgrad', k['wgrad_igemm']['ms'], 'loss', d['loss_after'])
except Exception as e: print('$tag', 'ERR', e)
"; tail -3 $o/$tag.err; }
b n3_read_a A=1
b n3_full_a UB_LIB_VARIANT=fullwait
b n3_read_b A=1
b n3_full_b UB_LIB_VARIANT=fullwait
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_layers_gpu.py tests/test_blocks_gpu.py -m gpu -q -x > $o/n3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $o/n3_pytest_gpu.log
