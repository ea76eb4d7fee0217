mkdir -p gpurun_out
timeout 600 python scratch/lin_dbg.py > gpurun_out/lin_dbg.log 2>&1; cat gpurun_out/lin_dbg.log | tail -20
