cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/bench_e2e_sub.py 8192 8192,4096,2048 > gpurun_out/r2_e2e_sub.log 2>&1; cat gpurun_out/r2_e2e_sub.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --sweep-particles 400000 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -3 gpurun_out/r2_bench_n2.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_n2.json'))
print(j['value'], j['e2e']['value'], j['full_sweep'])
PY
