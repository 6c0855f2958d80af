cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/bench_score.py 131072 2 0 1 1 0 1 2>&1 | tail -1
for f in 0 32; do
export ABC_MMA_FLAGS=$f
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:abc_score_m -c 6 --csv --log-file gpurun_out/r2_score_mma_launches.csv python scripts/bench_score.py 131072 2 0 1 0 0 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_score_mma_launches.csv') if not l.startswith('=='))]
h=rows[0]
for r in rows[-3:]:
    print("flags $f (32 = stage 3 writes no background)", r[h.index('Kernel Name')][:40], r[-1], r[-2])
PY
done
