# ncu --set full captures of the kernels the round-2 numbers rest on (one GPU; reports come back in gpurun_out/)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:onesided_round_resident -s 3000 -c 1 -o gpurun_out/prof_jacobi16_b2 -f python scratch/c2eig_b2.py > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:smallk_kernel -c 4 -o gpurun_out/prof_smallk -f python bench.py --warmup 3 --ncu-step > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:col2im_kernel -c 2 -o gpurun_out/prof_col2im -f python bench.py --warmup 3 --ncu-step > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 3 -c 1 -o gpurun_out/prof_gram_ts -f python profiles/run_gram.py > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
ls -la gpurun_out/*.ncu-rep
