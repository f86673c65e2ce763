#!/bin/bash
# GPU visit for the int8 digit-slice path: hardware convention test, parity tests, slices-vs-fp64 comparison.
TAG=${1:-oz}
mkdir -p gpurun_out
timeout 120 tools/bin/umma_test > gpurun_out/umma_test_$TAG.log 2>&1; echo "umma_test rc=$?"; tail -10 gpurun_out/umma_test_$TAG.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python tools/ozaki_check.py > gpurun_out/ozaki_small_$TAG.log 2>&1; echo "ozaki small rc=$?"; tail -12 gpurun_out/ozaki_small_$TAG.log
timeout 900 python tools/ozaki_check.py big > gpurun_out/ozaki_big_$TAG.log 2>&1; echo "ozaki big rc=$?"; tail -8 gpurun_out/ozaki_big_$TAG.log
