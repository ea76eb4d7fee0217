# bench lines on N GPUs of one box: N=$1, workloads $2...
mkdir -p gpurun_out
n=$1; shift
for w in "$@"; do
  steps=5; [ "$w" = c4 ] && steps=2; [ "$w" = c5 ] && steps=2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --workload $w --steps $steps --warmup 3 > gpurun_out/r02_bench_${w}_n$n.json 2> gpurun_out/bench_${w}_n$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_${w}_n$n.json').read().strip().splitlines()[-1]); print('$w', d['n_gpus'], d['value'], d['e2e']['value'], d['gram_assembly_ms_per_step'], d['eigensolver_ms_per_step'], d['schedule']['parallelism'][:80])" || tail -5 gpurun_out/bench_${w}_n$n.err
done
timeout 600 python -m pytest tests/test_dist_gpu.py -q 2>&1 | tail -2
