mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scratch/time_big.py
python bench.py --no-cpu-baseline --steps 5 --dtype f64 > gpurun_out/bench_c2_f64.json 2>gpurun_out/bench_c2_f64.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c2_f64.json')); print('c2 f64', d['value'], d['e2e']['value'], d['eigensolver'])"
