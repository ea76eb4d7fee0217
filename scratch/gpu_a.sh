mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do timeout 600 python bench.py --workload c3 --steps 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value']); [print(r) for r in d['kernels'][:3]]"; done
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('c2', d['value'], d['e2e']['value'])"
timeout 600 python bench.py --workload c5 --steps 2 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['value'], d['e2e']['value'])"
