#!/bin/bash
# compute-sanitizer passes over the GPU parity tests that exercise every kernel family (scoring both layouts + NaN paths,
# acceptance lists incl. the device sort, exact-math SSA, telegraph SSA modes 1/2 incl. corner schedules, adaptive burn-in is
# on by default, partition invariance, the pipelined abc_simulate_score).  Usage: bash scripts/run_sanitizer.sh [outfile]
OUT=${1:-gpurun_out/compute_sanitizer.txt}
SEL='score_bit_exact or special_values or accept_lists or exact_math or corner_cases or partition_invariant or simulate_score or empty_and_small'
: > "$OUT"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > /tmp/san_$tool.log 2>&1
  echo "$tool: exit $? ; $(grep -E 'passed|failed' /tmp/san_$tool.log | tail -1)" >> "$OUT"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san_$tool.log | tail -1 | sed "s/^/$tool: /" >> "$OUT"
  grep -E "Invalid|hazard|Race reported|Barrier error" /tmp/san_$tool.log | head -5 >> "$OUT"
done
cat "$OUT"
