mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/r02_bench_c2.json
for w in c1 c3 c4 c5; do timeout 900 python bench.py --workload $w --steps 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 200 gpurun_out/r02_bench_$w.json; done
timeout 600 python bench.py --dtype f64 --steps 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_f64.json 2> gpurun_out/bench_f64.err; tail -c 200 gpurun_out/r02_bench_c2_f64.json
timeout 300 python scratch/eig_time.py 320 640 1280 2560 5120 10240 > gpurun_out/r02_eig_time.log 2>&1; grep "^R=" gpurun_out/r02_eig_time.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --warmup 3 --ncu-step > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
python profiles/run_gram.py > gpurun_out/r02_run_gram.log 2>&1; tail -1 gpurun_out/r02_run_gram.log
