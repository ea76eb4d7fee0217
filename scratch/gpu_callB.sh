mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "bn1d or bn2d or bn3d" > gpurun_out/pytest_bn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_bn.log
tail -15 gpurun_out/pytest_bn.log
timeout 240 python scratch/cl_sweep.py > gpurun_out/cl_sweep2.log 2>&1; cat gpurun_out/cl_sweep2.log
