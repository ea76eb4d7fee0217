mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wide_rot_cluster -s 6 -c 1 -o gpurun_out/prof_wrotc -f python scratch/eig_time.py 5120 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 3 -c 1 -o gpurun_out/prof_gram_ts -f python profiles/run_gram.py > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
