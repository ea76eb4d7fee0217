mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --cpu-budget-s 40 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json | head -c 300
python bench.py --workload c1 --cpu-budget-s 10 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --workload c3 --cpu-budget-s 30 --steps 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 300 python bench.py --workload c4 --steps 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python bench.py --warmup 3 --ncu-step > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_c2.csv
for f in gpurun_out/bench_c*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['e2e']['value'], d.get('roofline',{}) and d['roofline'].get('frac'), d.get('eigensolver'))"; done
