cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score_mma.py -m gpu -x -q 2>&1 | tail -30
for mma in 0 1; do
  python scripts/bench_score.py 131072 2 0 1 1 0 $mma 2>&1 | tail -1
  python scripts/bench_score.py 131072 2 0 1 0 0 $mma 2>&1 | tail -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:abc_score -c 40 --csv --log-file gpurun_out/r2_score_mma_launches.csv python scripts/bench_score.py 131072 2 0 1 0 0 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_score_mma_launches.csv') if not l.startswith('=='))]
h=rows[0]
for r in rows[-4:]:
    print(r[h.index('Kernel Name')][:60], r[h.index('Grid Size')], r[-1], r[-2])
PY
