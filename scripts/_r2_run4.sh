cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ABC_BENCH_DEBUG=1 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --sweep-particles 0 --no-cpu > gpurun_out/r2_bench_async.json 2> gpurun_out/r2_bench_async.err; grep "e2e 1\]\|e2e 2\]" gpurun_out/r2_bench_async.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2_bench_async.json').read().splitlines() if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['share_of_step'])
PY
