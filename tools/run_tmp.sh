o=gpurun_out; mkdir -p $o
b() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --steps 30 --warmup 10 > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), round(d['value'],1))
except Exception as e: print('$tag', 'ERR', e)
"; }
b1() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/$tag.json 2> $o/$tag.err; python -c "
import json
d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4))"; }
b1 s15_1gpu UB_DEBUG_BUCKETS=1
grep bucket $o/s15_1gpu.err | head -14
b1 s15_1gpu_nolc UB_LEVEL_CUTS=0
b s15_2gpu A=1
b s15_2gpu_nolc UB_LEVEL_CUTS=0
b s15_2gpu_cta8 NCCL_MAX_CTAS=8
b s15_2gpu_cta16 NCCL_MAX_CTAS=16
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 tools/timeline.py 32 $o/s15_timeline_2gpu.tsv > $o/s15_timeline_2gpu.txt 2>&1
grep -A14 "NCCL kernels" $o/s15_timeline_2gpu.txt | head -16; grep "step span" $o/s15_timeline_2gpu.txt
