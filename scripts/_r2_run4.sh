cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "data_summary" > gpurun_out/r2_pytest_new.log 2>&1; tail -25 gpurun_out/r2_pytest_new.log
