#!/bin/bash
# GPU visit for the int8 digit-slice path: hardware convention test, slices-vs-fp64 error / timing tables.
# usage: tools/gpu_ozaki.sh <tag> [small|big|cfg3|cfg3e ...]
TAG=${1:-oz}; shift
mkdir -p gpurun_out
timeout 120 tools/bin/umma_test > gpurun_out/umma_test_$TAG.log 2>&1; echo "umma_test rc=$?"; tail -3 gpurun_out/umma_test_$TAG.log
for WHAT in ${@:-small big}; do
  ARG=$WHAT; [ "$WHAT" = small ] && ARG=""
  timeout 900 python tools/ozaki_check.py $ARG > gpurun_out/ozaki_${WHAT}_$TAG.log 2>&1; echo "ozaki $WHAT rc=$?"; grep "refine=1" gpurun_out/ozaki_${WHAT}_$TAG.log | cut -c1-220
done
