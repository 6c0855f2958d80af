cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
for m in 1 3 5; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abc_tele -s 1 -c 1 -f -o gpurun_out/r2_tele_m$m python scripts/bench_ssa.py 4096 $m 96 10 prior 2 2 > gpurun_out/r2_ncu_m$m.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --sweep-particles 0 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
( time timeout 1200 python bench.py --gpus 1 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2> gpurun_out/r2_bench_n1.time
tail -2 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.time
