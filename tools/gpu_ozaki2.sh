#!/bin/bash
# quick GPU visit for the int8 digit-slice path only: slices-vs-fp64 comparison at small and 32^3 sizes
TAG=${1:-oz}
mkdir -p gpurun_out
timeout 600 python tools/ozaki_check.py > gpurun_out/ozaki_small_$TAG.log 2>&1; echo "ozaki small rc=$?"; tail -12 gpurun_out/ozaki_small_$TAG.log
timeout 900 python tools/ozaki_check.py big > gpurun_out/ozaki_big_$TAG.log 2>&1; echo "ozaki big rc=$?"; tail -8 gpurun_out/ozaki_big_$TAG.log
