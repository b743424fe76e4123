o=gpurun_out; mkdir -p $o
timeout 100 python tools/run_reference_cuda.py 40 > $o/e1_ref_plain.txt 2>&1; tail -12 $o/e1_ref_plain.txt
timeout 200 python tools/run_reference_cuda.py 150 $o/e1_ref_launches.csv 5000 4000 > $o/e1_ref_ncu.txt 2>&1; tail -5 $o/e1_ref_ncu.txt
python tools/launch_summary.py $o/e1_ref_launches.csv > $o/e1_ref_launch_summary.txt 2>&1; head -40 $o/e1_ref_launch_summary.txt
