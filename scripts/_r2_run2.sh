cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/bench_ssa.py 32768 12345 96 10 prior 2 2 > gpurun_out/r2_bench_ssa_32k.log 2>&1
for m in 1 3 5; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abc_tele -s 1 -c 1 -o gpurun_out/r2_tele_m$m python scripts/bench_ssa.py 4096 $m 96 10 prior 2 2 > gpurun_out/r2_ncu_m$m.log 2>&1
done
cat gpurun_out/r2_bench_ssa_32k.log
