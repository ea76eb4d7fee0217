mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_c4_n2.err; head -c 700 gpurun_out/bench_c4_n2.json; echo
timeout 120 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_c2_quick.json 2> gpurun_out/bench_c2_quick.err; echo "rc=$?"; head -c 400 gpurun_out/bench_c2_quick.json
