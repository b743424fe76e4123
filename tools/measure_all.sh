#!/bin/bash
# Everything the round's profiles/ directory is built from, in one gpurun call (1 x B200):
#   bash tools/measure_all.sh <tag>          (outputs: gpurun_out/<tag>_*)
# 1. full GPU test suite  2. bench.py (default arguments)  3. ncu launch list of one eager step (device time per launch)
# 4. DRAM bytes + time of every tcgen05 launch of one step (roofline `traffic`)  5. per-op event times  6. conv microbench (BASELINE configs[1])  7. GEMM-kernel harness (CPU-checked)
# 8. ncu --set full of the hooked conv, wgrad and attention kernels (one ncu "use" per call: all runs below are ncu)
tag=${1:-final}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $o/${tag}_pytest_gpu.log
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
timeout 120 python tools/profile_step.py 3 > $o/${tag}_plain.log 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 892 -c 446 --csv \
    --log-file $o/${tag}_launches.csv python tools/profile_step.py 3 > $o/${tag}_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'igemm|attn_tc' -s 560 -c 280 --csv --log-file $o/${tag}_tc_traffic.csv python tools/profile_step.py 3 \
    > $o/${tag}_ncu_traffic.log 2>&1
timeout 120 python tools/profile_ops.py > $o/${tag}_ops.txt 2>&1
timeout 600 python tools/conv_bench.py --reps 20 --json $o/${tag}_conv_bench.json > $o/${tag}_conv_bench.txt 2>&1
timeout 200 unet.cu_b200/build/igemm_test > $o/${tag}_igemm_test.log 2>&1; echo "igemm_test rc=$?"
# step 3 of the tape order: 147 conv launches per step (57 forward, 90 backward); the last ones of a step are the
# 64x64 level's dgrad convs with the GroupNorm-backward hook
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'igemm_conv|igemm_rows' \
    -s 430 -c 6 -o $o/${tag}_prof_conv python tools/profile_step.py 3 > $o/${tag}_ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_wgrad_kernel -s 162 -c 4 \
    -o $o/${tag}_prof_wgrad python tools/profile_step.py 3 > $o/${tag}_ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 66 -c 4 \
    -o $o/${tag}_prof_attn python tools/profile_step.py 3 > $o/${tag}_ncu_attn.log 2>&1
tail -c 300 $o/${tag}_bench.json
