#!/bin/bash
# One GPU visit: parity tests, default bench, ncu launch list and one full capture of the top kernel.
# usage: tools/gpu_round.sh <tag> [workload]
TAG=${1:-r1}; WL=${2:-cfg2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py --workload $WL > gpurun_out/bench_${TAG}_$WL.json 2> gpurun_out/bench_${TAG}_$WL.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_$WL.json; tail -3 gpurun_out/bench_${TAG}_$WL.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_$WL.csv \
    python bench.py --workload $WL --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -c 1 -f -o gpurun_out/prof_project_${TAG}_$WL \
    python bench.py --workload $WL --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
