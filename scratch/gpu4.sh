mkdir -p gpurun_out
for n in 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_c2_n$n.json 2> gpurun_out/bench_c2_n$n.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_n$n.json').read().strip().splitlines()[-1]); print('c2', d['n_gpus'], d['value'], d['e2e'], d['gram_assembly_ms_per_step'], d['eigensolver_ms_per_step'], d['config']['parallelism']); [print(r) for r in d['kernels'][:4]]" || tail -5 gpurun_out/bench_c2_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 4 --workload c4 --steps 2 --warmup 3 > gpurun_out/bench_c4_n4.json 2> gpurun_out/bench_c4_n4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_n4.json').read().strip().splitlines()[-1]); print('c4', d['n_gpus'], d['value'], d['e2e'], d['config']['parallelism'])" || tail -5 gpurun_out/bench_c4_n4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --workload c5 --steps 2 --warmup 3 > gpurun_out/bench_c5_n4.json 2> gpurun_out/bench_c5_n4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c5_n4.json').read().strip().splitlines()[-1]); print('c5', d['n_gpus'], d['value'], d['e2e'], d['gram_assembly_ms_per_step'], d['config']['parallelism'])" || tail -5 gpurun_out/bench_c5_n4.err
