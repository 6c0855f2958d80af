cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python bench.py --gpus 1 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2> gpurun_out/r2_bench_n1.time
tail -3 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.time; head -c 600 gpurun_out/r2_bench_n1.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
