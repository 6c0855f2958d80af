cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score_mma.py -m gpu -x -q 2>&1 | tail -5
for mma in 0 1; do for lay in 2 1; do
python scripts/bench_score.py 131072 $lay 0 1 1 0 $mma 2>&1 | tail -1
done; done
python scripts/bench_score.py 524288 2 0 1 1 0 0 2>&1 | tail -1
python scripts/bench_score.py 524288 2 0 1 1 0 1 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:abc_score -c 40 --csv --log-file gpurun_out/r2_score_mma_launches.csv python scripts/bench_score.py 131072 2 0 1 0 0 1 > /dev/null 2>&1
for k in mma_filter mask_exact; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abc_score_$k -s 2 -c 1 -f -o gpurun_out/r2_score_$k python scripts/bench_score.py 131072 2 0 1 0 0 1 > gpurun_out/r2_ncu_$k.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_score_$k.ncu-rep > gpurun_out/r2_score_${k}_ncu_summary.csv
done
grep -i "tensor\|pipe_tc\|tmem\|gpu__time\|dram__bytes\|issue_active" gpurun_out/r2_score_mma_filter_ncu_summary.csv | head -20
ncu -i gpurun_out/r2_score_mma_filter.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if ('tensor' in a or 'pipe_tc' in a or 'tmem' in a.lower() or 'utc' in a.lower()) and c not in ('','0'): print(a,b,c)
" | head -30
