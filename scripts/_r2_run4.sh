cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "model_probs or recover_statistics or refuses or start_times or wrapper" > gpurun_out/r2_pytest_new.log 2>&1; tail -15 gpurun_out/r2_pytest_new.log
