cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score_mma.py -m gpu -x -q 2>&1 | tail -3
python scripts/bench_score.py 131072 2 0 1 1 0 1 2>&1 | tail -1
python scripts/bench_score.py 131072 1 0 1 1 0 1 2>&1 | tail -1
python scripts/bench_score.py 524288 2 0 1 1 0 1 2>&1 | tail -1
