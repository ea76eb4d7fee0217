mkdir -p gpurun_out
timeout 120 python scratch/gemm_lat.py 2>&1 | tail -8
echo "--- no tcgen05"; VVT_NO_TCGEN05=1 timeout 120 python scratch/gemm_lat.py 2>&1 | tail -8
echo "--- debug"; VVT_SYEVJ_DEBUG=1 timeout 120 python scratch/gemm_lat.py 2>&1 | grep "sweep 1:\|sweep 2:" | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python bench.py --warmup 3 --ncu-step > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_c2.csv
