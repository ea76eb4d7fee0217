timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "syevj" 2>&1 | grep -E "passed|failed|FAILED|Error" | head -5
VVT_SYEVJ_DEBUG=1 timeout 120 python scratch/one_syevj.py 1 2>&1 | grep "sweep 3:" | head -1 | sed 's/.*| tc apply/tc apply/' | cut -c1-200
