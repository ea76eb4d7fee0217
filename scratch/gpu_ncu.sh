mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "closures" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:onesided_round -s 300 -c 2 -f -o gpurun_out/r01_jacobi python scratch/one_syevj.py 1 > gpurun_out/ncu_jacobi.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 3 -c 1 -f -o gpurun_out/r01_gram_tc python profiles/run_gram.py > gpurun_out/ncu_gram.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:backtransform_dense -s 58 -c 1 -f -o gpurun_out/r01_backtransform python scratch/bt_bench.py > gpurun_out/ncu_bt.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chol_panel -s 0 -c 1 -f -o gpurun_out/r01_chol_panel python scratch/one_syevj.py 1 > gpurun_out/ncu_chol.log 2>&1
ls -la gpurun_out/*.ncu-rep
