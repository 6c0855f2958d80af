cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --sweep-particles 0 --no-cpu > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -2 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
txt=open('gpurun_out/r2_bench_n2.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['ms_per_step'], j['ms_busy_per_step_by_rank'])
PY
