mkdir -p gpurun_out
for v in 0 1 2 3; do VVT_WIDE_DESC=$v timeout 120 python scratch/wide_dbg.py 256 > gpurun_out/wide_dbg_$v.log 2>&1; cat gpurun_out/wide_dbg_$v.log | tail -12; done
