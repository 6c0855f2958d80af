cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash scripts/run_sanitizer.sh gpurun_out/r2_compute_sanitizer.txt
