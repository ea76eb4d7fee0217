mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_c2_n2.err; head -c 1200 gpurun_out/bench_c2_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference --cpu-budget-s 20 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "rc=$?"; cat gpurun_out/bench_ref_n2.json | head -c 600
