#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$1.log
for WL in cfg3e cfg3; do
timeout 1500 python bench.py --workload $WL --steps 2 --e2e-steps 1 > gpurun_out/bench_$1_$WL.json 2> gpurun_out/bench_$1_$WL.err; echo "bench $WL rc=$?"; cat gpurun_out/bench_$1_$WL.json; tail -3 gpurun_out/bench_$1_$WL.err
done
