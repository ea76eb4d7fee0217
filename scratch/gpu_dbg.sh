mkdir -p gpurun_out
for cl in 2 3; do echo "CL=$cl"; VVT_SYEVJ_CL=$cl python scratch/one_syevj.py 5; done > gpurun_out/syevj_cl.log 2>&1
VVT_SYEVJ_CL=3 VVT_SYEVJ_DEBUG=1 python scratch/one_syevj.py 1 2>&1 | grep "sweep 3" | tail -1 >> gpurun_out/syevj_cl.log
cat gpurun_out/syevj_cl.log
