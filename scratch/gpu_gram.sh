timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gram or gemm or conv or syevj" 2>&1 | tail -2
python profiles/run_gram.py
python scratch/one_syevj.py 5
python scratch/time_syevj.py 2>&1 | grep float32
