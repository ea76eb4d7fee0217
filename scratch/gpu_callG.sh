mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_c2_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_n2.json')); print('c2 n2', d['value'], d['e2e']['value'], d.get('scaling'), {k: d[k] for k in d if 'sharded' in k or 'replicated' in k or 'eigensolver_ms' in k})"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
echo "rc=$?"; tail -c 400 gpurun_out/bench_c4_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n2.json')); print('c4 n2', d['value'], d['e2e']['value'])"
