mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --cpu-budget-s 30 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --workload c1 --cpu-budget-s 8 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --workload c3 --cpu-budget-s 20 --steps 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 200 python bench.py --workload c4 --steps 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for f in gpurun_out/bench_c*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['e2e']['value'], d.get('roofline',{}) and d['roofline'].get('frac'), d.get('eigensolver'))"; done
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
