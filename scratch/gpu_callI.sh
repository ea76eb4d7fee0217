mkdir -p gpurun_out
timeout 60 python scratch/measure_peaks.py > gpurun_out/peaks_tf32_fp64.json 2>/dev/null; cat gpurun_out/peaks_tf32_fp64.json
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; echo "rc=$?"; wc -l gpurun_out/bench_c2_final.json; head -c 300 gpurun_out/bench_c2_final.json; echo
timeout 60 python bench.py --impl reference --steps 1 --warmup 1 --cpu-budget-s 5 --workload c1 2>/dev/null | head -c 200
