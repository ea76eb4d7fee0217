mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "wide_round or syevj" > gpurun_out/t_eig.log 2>&1; tail -4 gpurun_out/t_eig.log
timeout 300 python scratch/eig_time.py 5120 10240 > gpurun_out/eig_time.log 2>&1; grep "^R=" gpurun_out/eig_time.log
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:wide -c 90 --csv --log-file gpurun_out/ncu_wide.csv python scratch/eig_time.py 5120 > gpurun_out/ncu_wide.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/ncu_wide.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi: agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in agg.items(): print(k, len(v), 'mean', sum(v) / len(v), 'min', min(v), 'max', max(v), rows[hdr+1][ui])
PY
timeout 600 python bench.py --workload c4 --steps 2 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['eigensolver']); [print(r) for r in d['kernels'][:4]]"
