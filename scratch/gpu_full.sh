mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_c2_b.json 2>gpurun_out/bench_c2_b.err
python bench.py --workload c4 --no-cpu-baseline --steps 2 > gpurun_out/bench_c4_b.json 2>gpurun_out/bench_c4_b.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_c2_b.json','gpurun_out/bench_c4_b.json'):
    d=json.load(open(f))
    print(f, d['value'], d['e2e']['value'], d['eigensolver'])
    for k in d['kernels'][:6]: print(k)
PY
