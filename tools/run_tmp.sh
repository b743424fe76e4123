o=gpurun_out; mkdir -p $o
run() { tag=$1; np=$2; shift 2; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $np --steps 40 --warmup 10 > $o/$tag.json 2> $o/$tag.err; python -c "
import json
try:
    d=json.loads(open('$o/$tag.json').read().strip().splitlines()[-1]); print('$tag', 'gpus', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'img/s', round(d['value'],1))
except Exception as e: print('$tag', 'ERR', e)
"; grep -iE "error|fail" $o/$tag.err | head -2 | cut -c1-200; }
run n_cta8_tail1 8 NCCL_MAX_CTAS=8 UB_TAIL_NODE=1
run n_cta8_tail2 8 NCCL_MAX_CTAS=8 UB_TAIL_NODE=2
run n_cta8_tail2_ll128 8 NCCL_MAX_CTAS=8 UB_TAIL_NODE=2 NCCL_PROTO=LL128
run n_cta8_tail2_nvls 8 NCCL_MAX_CTAS=8 UB_TAIL_NODE=2 NCCL_ALGO=NVLS
run n_cta8_tail2_tree 8 NCCL_MAX_CTAS=8 UB_TAIL_NODE=2 NCCL_ALGO=Tree
run n_cta16_tail2 8 NCCL_MAX_CTAS=16 UB_TAIL_NODE=2
