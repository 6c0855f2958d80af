cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2_n8_gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu_from_one_process" > gpurun_out/r2_pytest_multi8.log 2>&1; tail -3 gpurun_out/r2_pytest_multi8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/check_multi_gpu.py 20000 > gpurun_out/r2_check_multi_n8.log 2>&1; tail -2 gpurun_out/r2_check_multi_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -2 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
txt=open('gpurun_out/r2_bench_n8.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'], j['full_sweep'])
PY
