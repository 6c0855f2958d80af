cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "start_times or adaptive or refuses or ssa_moments or hybrid_burnin_equals" > gpurun_out/r2_pytest_ssa.log 2>&1; tail -3 gpurun_out/r2_pytest_ssa.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"abc_window|abc_tele" -c 20 --csv --log-file gpurun_out/r2_win_launches.csv python scripts/bench_ssa.py 8192 12345 96 10 prior 2 2 > /dev/null 2>&1
grep -E "abc_window|abc_tele" gpurun_out/r2_win_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-
timeout 300 python scripts/bench_ssa.py 8192 12345 96 10 prior 2 2 | grep mode
