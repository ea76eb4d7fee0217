# First gpurun call of round 2: everything that was built after the last GPU minute of round 1.
#   gpurun --timeout 1500 -- 'bash scratch/gpu_round2_first.sh'
mkdir -p gpurun_out
# 1. the full GPU suite.  The tests added after the last full run (test_axpy, the branching fixture's cases,
#    the reference-run vectors, Linear closures, NTK) already ran green on a B200 on their own
#    (profiles/r01_pytest_gpu_added_tests.log); the conv handlers were refactored after that (Conv1d support),
#    so the full suite is the first thing to confirm.  Then add a 1-d fixture (Conv1d / MaxPool1d / AvgPool1d)
#    to tests/problems.py: it would be the first GPU case with non-square kernels and strides.
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# 2. smoke + the headline bench line
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json
# 3. the experimental wide stage of the eigensolver against the default path
timeout 300 python scratch/wide_check.py > gpurun_out/wide_default.log 2>&1; tail -4 gpurun_out/wide_default.log
VVT_SYEVJ_WIDE=1 VVT_SYEVJ_DEBUG=1 timeout 300 python scratch/wide_check.py > gpurun_out/wide_1.log 2>&1; grep "^R=" gpurun_out/wide_1.log
VVT_SYEVJ_WIDE=2 timeout 300 python scratch/wide_check.py > gpurun_out/wide_2.log 2>&1; grep "^R=" gpurun_out/wide_2.log
