cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for l in libabcb200 libabcb200_u23; do
  echo "== $l"
  ABCB200_LIB=$PWD/abc_inference_transcription_b200/$l.so python scripts/bench_ssa.py 8192 12345 2>&1 | tail -2
done
done
echo "== fixed burn-in"
for l in libabcb200 libabcb200_u23; do
  ABCB200_LIB=$PWD/abc_inference_transcription_b200/$l.so python scripts/bench_ssa.py 8192 12345 96 10 prior 2 0 2>&1 | tail -2
done
