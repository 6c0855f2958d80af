cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ode.py -m gpu -x -q > gpurun_out/r2_pytest_ode.log 2>&1; tail -8 gpurun_out/r2_pytest_ode.log
timeout 900 python scripts/run_equivalence.py 100000 96 gpurun_out/r2_equivalence_ssa_vs_ode_1e5.json > gpurun_out/r2_equiv.log 2>&1; tail -5 gpurun_out/r2_equiv.log | cut -c1-400
