#!/bin/bash
TAG=${1:-oz}
mkdir -p gpurun_out
timeout 600 python tools/ozaki_check.py > gpurun_out/ozaki_small_$TAG.log 2>&1; echo "ozaki small rc=$?"; tail -19 gpurun_out/ozaki_small_$TAG.log
timeout 900 python tools/ozaki_check.py big > gpurun_out/ozaki_big_$TAG.log 2>&1; echo "ozaki big rc=$?"; tail -13 gpurun_out/ozaki_big_$TAG.log
