#!/bin/bash
# One GPU visit: parity tests, smoke, default bench (+ reference arm), cfg3 bench, ncu launch list and full captures.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_cfg2.json; tail -3 gpurun_out/bench_${TAG}_cfg2.err
timeout 900 python bench.py --workload cfg3 --steps 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "bench cfg3 rc=$?"; cat gpurun_out/bench_${TAG}_cfg3.json; tail -3 gpurun_out/bench_${TAG}_cfg3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${TAG}_cfg2.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_project -c 1 -f -o gpurun_out/prof_project_${TAG}_cfg2 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full project rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -c 4 -f -o gpurun_out/prof_gemm_${TAG}_cfg2 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full2_$TAG.log 2>&1; echo "ncu full gemm rc=$?"
ls -la gpurun_out | tail -12
