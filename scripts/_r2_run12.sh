cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --gpus 1 > gpurun_out/r2_bench_n1_final.log 2>&1; tail -1 gpurun_out/r2_bench_n1_final.log > gpurun_out/r2_bench_n1_final.json; tail -c 1500 gpurun_out/r2_bench_n1_final.json
