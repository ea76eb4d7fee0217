mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -q > gpurun_out/t_dist.log 2>&1; tail -15 gpurun_out/t_dist.log
for w in c2 c5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --workload $w --steps 3 --warmup 3 > gpurun_out/bench_${w}_n2.json 2> gpurun_out/bench_${w}_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_n2.json').read().strip().splitlines()[-1]); print('$w', d['n_gpus'], d['value'], d['e2e'], d['gram_assembly_ms_per_step'], d['eigensolver_ms_per_step']); [print(r) for r in d['kernels'][:5]]" || tail -5 gpurun_out/bench_${w}_n2.err
done
