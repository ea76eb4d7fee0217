mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_bench_size_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests/test_bench_size_gpu.py -q -s > gpurun_out/pytest_bench_size.log 2>&1; tail -30 gpurun_out/pytest_bench_size.log
timeout 600 python bench.py --steps 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c4 --steps 2 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/bench_c4.json
