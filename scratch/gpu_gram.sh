timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gram or gemm or conv or syevj" 2>&1 | tail -2
python profiles/run_gram.py
python profiles/run_gram.py 1280 55296
python profiles/run_gram.py 5120 4096
