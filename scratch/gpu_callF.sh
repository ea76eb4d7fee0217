mkdir -p gpurun_out
timeout 400 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "rc=$?"; tail -3 gpurun_out/bench_c5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print('c5', d['value'], d['e2e']['value'], d.get('eigensolver')); print([(k['kernel'], k['ms_per_step']) for k in d['kernels'][:6]])"
