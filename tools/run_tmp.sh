o=gpurun_out; mkdir -p $o
# n1: wgrad epilogue through TMA reduce-add boxes (UB_WGRAD_TMA_RED=0 = the per-lane REDs it replaces)
timeout 200 unet.cu_b200/build/igemm_test > $o/n1_igemm_test.log 2>&1; echo "igemm_test rc=$?"; grep -A3 "^wgrad" $o/n1_igemm_test.log | tail -60; tail -2 $o/n1_igemm_test.log
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); k=d['kernel_classes']; print('$tag', 'ms', round(d['ms_per_step'],4), 'wgrad', k['wgrad_igemm']['ms'], k['wgrad_igemm']['tflops'], 'loss', d['loss_after'])
except Exception as e: print('$tag', 'ERR', e)
"; tail -3 $o/$tag.err; }
b n1_tmared A=1
b n1_lanered UB_WGRAD_TMA_RED=0
b n1_tmared_k512 UB_WGRAD_MIN_KPIX=512
b n1_tmared2 A=1
timeout 600 python -m pytest tests -m gpu -q -x -rs > $o/n1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $o/n1_pytest_gpu.log
