timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "syevj" 2>&1 | tail -2
VVT_SYEVJ_DEBUG=1 python scratch/one_syevj.py 1 2>&1 | grep "sweep 3:" | head -1 | sed 's/.*chol/chol/'
python scratch/one_syevj.py 5
python scratch/time_syevj.py 2>&1 | tail -8
