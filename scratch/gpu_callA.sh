mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu -k "center_rows or gram_extension_hooks or additional" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
tail -25 gpurun_out/pytest_new.log
timeout 150 python scratch/cl_sweep.py > gpurun_out/cl_sweep.log 2>&1; cat gpurun_out/cl_sweep.log
