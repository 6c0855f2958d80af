#!/bin/bash
# One pass over the GPU evidence of a build (run on a B200 box from the repo root; outputs under gpurun_out/):
#   bash scripts/gpu_validate.sh            tests + default bench line
#   bash scripts/gpu_validate.sh score      + scoring A/B (default path vs tensor-core filter), launch list, ncu --set full captures
#   bash scripts/gpu_validate.sh sanitizer  + compute-sanitizer passes (scripts/run_sanitizer.sh and the tensor-core scoring tests)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --gpus 1 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log > gpurun_out/bench_n1.json
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_n1.json').read())
print({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, j['e2e']['value'], j['roofline']['frac'],
      j['roofline_score']['ms_per_launch'], j['roofline_score']['frac'], j['roofline_score']['tensor_core_filter'])
PY
for what in "$@"; do
  if [ "$what" = score ]; then
    for mma in 0 1; do for lay in 2 1; do python scripts/bench_score.py 131072 $lay 0 1 1 0 $mma 2>&1 | tail -1; done; done | tee gpurun_out/score_mma_ab.txt
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:abc_score -c 40 --csv --log-file gpurun_out/score_mma_launches.csv python scripts/bench_score.py 131072 2 0 1 0 0 1 > /dev/null 2>&1
    for k in mma_filter mask_exact; do
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:abc_score_$k -s 2 -c 1 -f -o gpurun_out/score_$k python scripts/bench_score.py 131072 2 0 1 0 0 1 > gpurun_out/ncu_$k.log 2>&1
      python scripts/ncu_summary.py gpurun_out/score_$k.ncu-rep > gpurun_out/score_${k}_ncu_summary.csv
    done
  elif [ "$what" = sanitizer ]; then
    bash scripts/run_sanitizer.sh gpurun_out/compute_sanitizer.txt
    for tool in memcheck synccheck; do
      timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_score_mma.py -m gpu -q -x \
        -k 'accumulators or bit_exact_both_layouts or special_values or signed_and_degenerate or near_matches or empty_and_small' 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | sed "s/^/$tool (tensor-core scoring): /"
    done | tee gpurun_out/compute_sanitizer_mma.txt
  fi
done
