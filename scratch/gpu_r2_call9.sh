mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_bench_size_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests/test_bench_size_gpu.py -q -s > gpurun_out/pytest_bench_size.log 2>&1; grep -n "kink\|fell back\|passed\|failed\|^E  " gpurun_out/pytest_bench_size.log | head -20
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['eigensolver']); [print(r) for r in d['kernels'][:4]]"
timeout 600 python bench.py --workload c1 --steps 10 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['eigensolver']); [print(r) for r in d['kernels'][:4]]"
timeout 900 python bench.py --workload c5 --steps 2 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['eigensolver']); [print(r) for r in d['kernels'][:6]]"
