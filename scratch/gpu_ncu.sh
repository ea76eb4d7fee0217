mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wide_tc_kernel -s 6 -c 1 -o gpurun_out/prof_wgram -f python scratch/eig_time.py 5120 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:wide_apply -s 6 -c 1 -o gpurun_out/prof_wapply -f python scratch/eig_time.py 5120 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on -k regex:wide_rot -s 6 -c 1 -o gpurun_out/prof_wrot -f python scratch/eig_time.py 5120 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
