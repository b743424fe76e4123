#!/bin/bash
# Everything the round's profiles/ directory is built from, in one gpurun call (1 x B200):
#   bash tools/measure_all.sh <tag>          (outputs: gpurun_out/<tag>_*)
# 1. full GPU test suite  2. bench.py (default arguments) + the reference arm  3. ncu launch list of one eager step
# (device time per launch)  4. DRAM bytes + time of every tcgen05 launch of one step (roofline `traffic`)  5. per-op event
# times  6. device timeline of the captured step (CUPTI)  7. conv microbench (BASELINE configs[1])  8. GEMM-kernel harness
# (CPU-checked)  9. ncu --set full of the hooked conv, wgrad, attention kernels and of the HBM-bound kernels (GroupNorm,
# data movement, AdamW, weight re-pack).  All ncu runs of a call count as one ncu use; each follows a plain run of the same command.
# SHORT=1 bash tools/measure_all.sh <tag>: only 1-5 and the wgrad ncu capture (what changes when one kernel changed).
tag=${1:-final}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -rs > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $o/${tag}_pytest_gpu.log
timeout 900 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
if [ -z "$SHORT" ]; then timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"; fi
timeout 120 python tools/profile_step.py 3 > $o/${tag}_plain.log 2>&1 || { echo "plain run failed"; cat $o/${tag}_plain.log; exit 1; }
n=$(grep -o "launches/step [0-9]*" $o/${tag}_plain.log | awk '{print $2}'); echo "launches per eager step: $n"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * n)) -c $n --csv \
    --log-file $o/${tag}_launches.csv python tools/profile_step.py 3 > $o/${tag}_ncu_launches.log 2>&1
ntc=$(python tools/launch_summary.py $o/${tag}_launches.csv --count 'igemm|attn_tc' 2>/dev/null || echo 280)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'igemm|attn_tc' -s $((2 * ntc)) -c $ntc --csv --log-file $o/${tag}_tc_traffic.csv python tools/profile_step.py 3 \
    > $o/${tag}_ncu_traffic.log 2>&1
timeout 120 python tools/profile_ops.py > $o/${tag}_ops.txt 2>&1
timeout 120 python tools/timeline.py 32 $o/${tag}_timeline.tsv > $o/${tag}_timeline.txt 2>&1
nwg=$(python tools/launch_summary.py $o/${tag}_launches.csv --count 'igemm_wgrad_kernel' 2>/dev/null || echo 81)
if [ -n "$SHORT" ]; then
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_wgrad_kernel -s $((2 * nwg)) -c 4 \
        -o $o/${tag}_prof_wgrad python tools/profile_step.py 3 > $o/${tag}_ncu_wgrad.log 2>&1
    tail -c 300 $o/${tag}_bench.json
    exit 0
fi
timeout 600 python tools/conv_bench.py --reps 20 --json $o/${tag}_conv_bench.json > $o/${tag}_conv_bench.txt 2>&1
timeout 200 unet.cu_b200/build/igemm_test > $o/${tag}_igemm_test.log 2>&1; echo "igemm_test rc=$?"
# the last conv launches of a step are the 64x64 level's dgrad convs with the GroupNorm-backward hook
nconv=$(python tools/launch_summary.py $o/${tag}_launches.csv --count 'igemm_conv|igemm_rows' 2>/dev/null || echo 147)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'igemm_conv|igemm_rows' \
    -s $((3 * nconv - 10)) -c 6 -o $o/${tag}_prof_conv python tools/profile_step.py 3 > $o/${tag}_ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_wgrad_kernel -s $((2 * nwg)) -c 4 \
    -o $o/${tag}_prof_wgrad python tools/profile_step.py 3 > $o/${tag}_ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 66 -c 4 \
    -o $o/${tag}_prof_attn python tools/profile_step.py 3 > $o/${tag}_ncu_attn.log 2>&1
# HBM-bound kernels: every launch of one step (third of three), the metrics the summary needs
nhbm=$(python tools/launch_summary.py $o/${tag}_launches.csv --count 'gn_apply|gn_bwd_apply|gn_stats|gn_bwd_stats|concat2|add2_kernel|adamw_kernel|pack_weights|avgpool2|upsample2' 2>/dev/null || echo 160)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'gn_apply|gn_bwd_apply|gn_stats|gn_bwd_stats|concat2|add2_kernel|adamw_kernel|pack_weights|avgpool2|upsample2' \
    -s $((2 * nhbm)) -c $nhbm --csv --log-file $o/${tag}_hbm_kernels.csv python tools/profile_step.py 3 > $o/${tag}_ncu_hbm.log 2>&1
tail -c 300 $o/${tag}_bench.json
