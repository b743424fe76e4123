o=gpurun_out; mkdir -p $o
timeout 300 python -m pytest tests/test_attention_nhwc_gpu.py tests/test_blocks_gpu.py -q -x > $o/d1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/d1_pytest.log
UB_ATTN_HSPLIT=7 timeout 300 python -m pytest tests/test_attention_nhwc_gpu.py -q -x > $o/d1_pytest7.log 2>&1; echo "pytest hsplit7 rc=$?"; tail -3 $o/d1_pytest7.log
UB_CONV_MIN_BN=32 timeout 300 python -m pytest tests/test_trainer_gpu.py -q -x -k "B4 or other_configs" > $o/d1_pytest_bn32.log 2>&1; echo "pytest minbn32 rc=$?"; tail -3 $o/d1_pytest_bn32.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); k=d['kernel_classes']; print('$tag', 'ms', round(d['ms_per_step'],4), 'conv', k['conv_igemm']['ms'], 'attn', k['attention']['ms'])
except Exception as e: print('$tag', 'ERR', e)
"; }
run d1_h1 UB_ATTN_HSPLIT=1
run d1_h0 UB_ATTN_HSPLIT=0
run d1_h7 UB_ATTN_HSPLIT=7
run d1_h5 UB_ATTN_HSPLIT=5
run d1_h1_bn32 UB_ATTN_HSPLIT=1 UB_CONV_MIN_BN=32
run d1_h0b UB_ATTN_HSPLIT=0
run d1_h1b UB_ATTN_HSPLIT=1
