mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "solve_queue" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err; tail -3 gpurun_out/b_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_c2.json').read().strip().splitlines()[-1])
print('c2', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['sequential_solves'])
print(d['eigensolver'])
for k in d['kernels'][:6]: print(' ', k['kernel'], k['ms_per_step'])
PY
