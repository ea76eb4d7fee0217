mkdir -p gpurun_out
for v in 0 1 4; do VVT_WIDE_DESC=$v timeout 120 python scratch/wide_dbg.py 256 > gpurun_out/wide_dbg_$v.log 2>&1; cat gpurun_out/wide_dbg_$v.log | tail -12; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "wide_round" > gpurun_out/t_wide.log 2>&1; tail -5 gpurun_out/t_wide.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "syevj or symeig or mixed or newton" > gpurun_out/t_eig.log 2>&1; tail -12 gpurun_out/t_eig.log
VVT_SYEVJ_DEBUG=1 timeout 300 python scratch/eig_time.py 1280 2560 5120 > gpurun_out/eig_time.log 2>&1; grep "^R=" gpurun_out/eig_time.log
