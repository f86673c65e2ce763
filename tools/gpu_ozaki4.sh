#!/bin/bash
TAG=${1:-oz}
mkdir -p gpurun_out
timeout 600 python tools/ozaki_check.py > gpurun_out/ozaki_small_$TAG.log 2>&1; echo "ozaki small rc=$?"; tail -12 gpurun_out/ozaki_small_$TAG.log
timeout 900 python tools/ozaki_check.py big > gpurun_out/ozaki_big_$TAG.log 2>&1; echo "ozaki big rc=$?"; tail -8 gpurun_out/ozaki_big_$TAG.log
timeout 1200 python -m pytest tests/test_gpu_int8.py -m gpu -x -q > gpurun_out/pytest_int8_$TAG.log 2>&1; echo "pytest int8 rc=$?"; tail -15 gpurun_out/pytest_int8_$TAG.log
