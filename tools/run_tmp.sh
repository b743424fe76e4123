o=gpurun_out; mkdir -p $o
# n2: second TMA producer warp in the wgrad kernel (UB_WGRAD_NPROD=1 = one producer)
timeout 200 unet.cu_b200/build/igemm_test > $o/n2_igemm_test.log 2>&1; echo "igemm_test rc=$?"; grep -A3 "^wgrad" $o/n2_igemm_test.log | grep -v "^--" | tail -40; tail -1 $o/n2_igemm_test.log
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); k=d['kernel_classes']; print('$tag', 'ms', round(d['ms_per_step'],4), 'wgrad', k['wgrad_igemm']['ms'], k['wgrad_igemm']['tflops'], 'loss', d['loss_after'])
except Exception as e: print('$tag', 'ERR', e)
"; tail -3 $o/$tag.err; }
b n2_prod2_a A=1
b n2_prod1_a UB_WGRAD_NPROD=1
b n2_prod2_b A=1
b n2_prod1_b UB_WGRAD_NPROD=1
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_layers_gpu.py -m gpu -q -x > $o/n2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $o/n2_pytest_gpu.log
