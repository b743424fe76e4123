o=gpurun_out
for m in 0 1 2; do for kb in 128 200; do
echo "== mode $m kb $kb"; UB_TRACE_MODE=$m UB_CONV2_SMEM_KB=$kb timeout 60 unet.cu_b200/build/igemm_trace shape 32 8 8 256 256 0 20 2>&1 | grep -E "first stage|last MMA|last TMA|accumulator|conv B32"
done; done > $o/s19_trace.txt 2>&1
echo "== nacc1"; UB_CONV_NACC=1 timeout 60 unet.cu_b200/build/igemm_trace shape 32 8 8 256 256 0 20 2>&1 | grep -E "first stage|last MMA|last TMA|conv B32" >> $o/s19_trace.txt
echo "== BN128"; UB_CONV_FORCE_BN=128 timeout 60 unet.cu_b200/build/igemm_trace shape 32 8 8 256 256 0 20 2>&1 | grep -E "first stage|last MMA|last TMA|conv B32" >> $o/s19_trace.txt
echo "== BN32"; UB_CONV_FORCE_BN=32 timeout 60 unet.cu_b200/build/igemm_trace shape 32 8 8 256 256 0 20 2>&1 | grep -E "first stage|last MMA|last TMA|conv B32" >> $o/s19_trace.txt
