mkdir -p gpurun_out
timeout 200 python scratch/cl_check.py > gpurun_out/cl_check.log 2>&1; cat gpurun_out/cl_check.log
echo "--- NOPAD"; VVT_SYEVJ_NOPAD=1 timeout 100 python scratch/cl_check.py small 2>&1 | tail -4
echo "--- PAD"; timeout 100 python scratch/cl_check.py small 2>&1 | tail -4
timeout 200 python bench.py --workload c4 --steps 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['e2e']['value'], d.get('eigensolver'))"
