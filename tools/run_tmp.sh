o=gpurun_out; mkdir -p $o
timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x -k "flip or B4 or golden" > $o/s1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/s1_pytest.log
for i in 1 2; do timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/s1_default$i.json 2> $o/s1_default$i.err; done
UB_EPI_MMA=1 timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x -k "B4 or golden" > $o/s1_pytest_epimma.log 2>&1; echo "epimma pytest rc=$?"; tail -3 $o/s1_pytest_epimma.log
UB_EPI_MMA=1 timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/s1_epimma.json 2> $o/s1_epimma.err
UB_CONV_NACC=4 timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x -k "B4 or golden" > $o/s1_pytest_nacc4.log 2>&1; echo "nacc4 pytest rc=$?"; tail -3 $o/s1_pytest_nacc4.log
UB_CONV_NACC=4 timeout 200 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-reference-cuda > $o/s1_nacc4.json 2> $o/s1_nacc4.err
for f in s1_default1 s1_default2 s1_epimma s1_nacc4; do python -c "
import json,sys
try:
    d=json.loads(open('$o/$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('$f', 'ERR', e)
"; done
