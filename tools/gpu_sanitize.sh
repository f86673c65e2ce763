#!/bin/bash
# compute-sanitizer over tiny inversions of every kernel family (SURVEY section 5 / VERDICT r1 item 8): memcheck, racecheck,
# synccheck, initcheck.  One GPU.  Logs -> gpurun_out/sanitize_<tool>_<tag>.log; copy the summaries to profiles/.
# usage: tools/gpu_sanitize.sh <tag> [tools...]
TAG=${1:-r2}; shift
TOOLS=${@:-memcheck racecheck synccheck initcheck}
mkdir -p gpurun_out
for T in $TOOLS; do
  EXTRA=""
  [ "$T" = "memcheck" ] && EXTRA="--leak-check full"
  [ "$T" = "racecheck" ] && EXTRA="--racecheck-report all"
  timeout 600 compute-sanitizer --tool $T $EXTRA --error-exitcode 9 --print-limit 40 \
      python tools/sanitize_cases.py > gpurun_out/sanitize_${T}_${TAG}.log 2>&1
  echo "sanitize $T rc=$?"; grep -c SANITIZE_CASE_OK gpurun_out/sanitize_${T}_${TAG}.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY" gpurun_out/sanitize_${T}_${TAG}.log | tail -3
done
