mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "syevj or symeig" > gpurun_out/t_eig.log 2>&1; tail -4 gpurun_out/t_eig.log
timeout 300 python scratch/c2eig.py 2>&1 | tail -1
VVT_SYEVJ_NOPERSIST=1 timeout 300 python scratch/c2eig.py 2>&1 | tail -1
timeout 300 python scratch/eig_time.py 320 640 1280 > gpurun_out/eig_time.log 2>&1; grep "^R=" gpurun_out/eig_time.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['eigensolver']); [print(r) for r in d['kernels'][:4]]"
timeout 600 python bench.py --workload c1 --steps 10 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['eigensolver']); [print(r) for r in d['kernels'][:3]]"
