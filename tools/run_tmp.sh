o=gpurun_out; mkdir -p $o
timeout 250 python -m pytest tests/test_dp_gpu.py -q -x > $o/g1_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -3 $o/g1_pytest_dp.log
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 40 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/g1_bench_2gpu.json 2> $o/g1_bench_2gpu.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$o/g1_bench_2gpu.json').read().strip().splitlines()[-1]); print('2gpu ms', round(d['ms_per_step'],4), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
