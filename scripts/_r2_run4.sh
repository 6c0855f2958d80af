cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_score_sweep.log
for cfg in "2 0 1 1 0 1" "2 0 1 1 0 0" "2 0 1 1 4 1" "2 0 1 0 0 1"; do timeout 120 python scripts/bench_score.py 131072 $cfg >> gpurun_out/r2_score_sweep.log 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:abc_score3 -c 40 --csv --log-file gpurun_out/r2_score_launches.csv python scripts/bench_score.py 131072 2 0 1 0 0 1 > /dev/null 2>&1
grep "n=" gpurun_out/r2_score_sweep.log; tail -8 gpurun_out/r2_score_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-50,120-
