mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "wide_round" > gpurun_out/t_wide.log 2>&1; tail -15 gpurun_out/t_wide.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "syevj or symeig or mixed or newton" > gpurun_out/t_eig.log 2>&1; tail -25 gpurun_out/t_eig.log
VVT_SYEVJ_DEBUG=1 timeout 300 python scratch/eig_time.py 320 1280 2560 5120 > gpurun_out/eig_time.log 2>&1; grep "^R=" gpurun_out/eig_time.log
