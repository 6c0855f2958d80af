cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ssa or simulate or multi_context" > gpurun_out/r2_pytest_ssa.log 2>&1; tail -3 gpurun_out/r2_pytest_ssa.log
rm -f gpurun_out/r2_bench_ssa.log
for ad in 2 1 0; do timeout 300 python scripts/bench_ssa.py 8192 12345 96 10 prior 2 $ad >> gpurun_out/r2_bench_ssa.log 2>&1; done
timeout 300 python scripts/bench_ssa.py 8192 45 96 10 corner 2 2 >> gpurun_out/r2_bench_ssa.log 2>&1
grep -E "mode|m=" gpurun_out/r2_bench_ssa.log
