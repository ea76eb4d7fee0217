mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "wide or two_level or batched" 2>&1 | tail -12
timeout 300 python scratch/eig_time.py 5120 10240 2>&1 | grep "^R="
VVT_WIDE_FULL_GRAM=1 timeout 300 python scratch/eig_time.py 5120 2>&1 | grep "^R="
