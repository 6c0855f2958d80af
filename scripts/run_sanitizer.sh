#!/bin/bash
# compute-sanitizer passes over the GPU parity tests that exercise every kernel family (scoring both layouts + NaN paths,
# acceptance lists incl. the device sort, exact-math SSA, telegraph SSA (product sampler, start-time rule) incl. corner schedules
# and the refusal path, mode 1, partition invariance, the pipelined abc_simulate_score, the model-probability bootstrap, the
# single-device multi context).  racecheck runs with 32 cells per read-out on a reduced selection so that it completes.
# Usage: bash scripts/run_sanitizer.sh [outfile]
OUT=${1:-gpurun_out/compute_sanitizer.txt}
SEL='score_bit_exact or special_values or accept_lists or exact_math or corner_cases or partition_invariant or simulate_score or empty_and_small or start_times or refuses or model_probs_on_device or multi_context'
RSEL='score_bit_exact or accept_lists or exact_math or partition_invariant or refuses or model_probs_on_device or multi_context or simulate_statistics_equal'
: > "$OUT"
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > /tmp/san_$tool.log 2>&1
  echo "$tool: exit $? ; $(grep -E 'passed|failed' /tmp/san_$tool.log | tail -1)" >> "$OUT"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san_$tool.log | tail -1 | sed "s/^/$tool: /" >> "$OUT"
  grep -E "Invalid|hazard|Race reported|Barrier error" /tmp/san_$tool.log | head -5 >> "$OUT"
done
tool=racecheck
timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$RSEL" > /tmp/san_$tool.log 2>&1
echo "$tool: exit $? ; $(grep -E 'passed|failed' /tmp/san_$tool.log | tail -1)" >> "$OUT"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san_$tool.log | tail -1 | sed "s/^/$tool: /" >> "$OUT"
grep -E "Invalid|hazard|Race reported|Barrier error" /tmp/san_$tool.log | head -5 >> "$OUT"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > /tmp/san_smoke.log 2>&1
echo "racecheck smoke(): exit $? ; $(grep -E 'RACECHECK SUMMARY|smoke\] ok' /tmp/san_smoke.log | tr '\n' ' ')" >> "$OUT"
cat "$OUT"
