timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pools or conv" 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_c2_b.json 2>gpurun_out/bench_c2_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2_b.json'))
print(d['value'], d['e2e']['value'])
for k in d['kernels'][:8]: print(k)
PY
