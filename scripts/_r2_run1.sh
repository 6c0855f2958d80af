cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_bench_ssa_var.log
for v in c4; do echo "== variant $v" >> gpurun_out/r2_bench_ssa_var.log; ABCB200_LIB=$GRAFT_REPO_ROOT/build/libabcb200_$v.so timeout 300 python scripts/bench_ssa.py 8192 12345 96 10 prior 2 2 >> gpurun_out/r2_bench_ssa_var.log 2>&1; done
echo "== default" >> gpurun_out/r2_bench_ssa_var.log; timeout 300 python scripts/bench_ssa.py 8192 12345 96 10 prior 2 2 >> gpurun_out/r2_bench_ssa_var.log 2>&1
grep -E "==|mode" gpurun_out/r2_bench_ssa_var.log
