timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "syevj" 2>&1 | tail -3
VVT_SYEVJ_DEBUG=1 timeout 120 python scratch/one_syevj.py 1 2>&1 | grep "sweep 3:" | head -1 | cut -c1-330
timeout 120 python scratch/one_syevj.py 5
timeout 200 python scratch/time_syevj.py 2>&1 | tail -8
