o=gpurun_out; mkdir -p $o
timeout 260 python tools/run_reference_cuda.py 110 $o/m1_ref_launches.csv 6000 6000 > $o/m1_ref_ncu.txt 2>&1; tail -6 $o/m1_ref_ncu.txt
ls -la $o/m1_ref_launches.csv
python tools/launch_summary.py $o/m1_ref_launches.csv > $o/m1_ref_launch_summary.txt 2>&1; head -30 $o/m1_ref_launch_summary.txt
