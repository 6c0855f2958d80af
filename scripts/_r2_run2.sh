cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abc_tele -s 1 -c 1 -o gpurun_out/r2_tele_corner python scripts/bench_ssa.py 8192 4 96 10 corner 2 2 > gpurun_out/r2_ncu_corner.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abc_tele -s 1 -c 1 -f -o gpurun_out/r2_tele_m1 python scripts/bench_ssa.py 4096 1 96 10 prior 2 2 > gpurun_out/r2_ncu_m1.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "start_times or refuses" > gpurun_out/r2_pytest_ssa2.log 2>&1; tail -3 gpurun_out/r2_pytest_ssa2.log
