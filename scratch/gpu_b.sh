timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "conv3d_and or one_dimensional" 2>&1 | tail -15
