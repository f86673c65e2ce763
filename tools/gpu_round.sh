#!/bin/bash
# One GPU visit: parity tests, smoke, default bench (+ CPU baseline), cfg3 bench, ncu launch lists and full captures.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_${TAG}_cfg2.json; tail -3 gpurun_out/bench_${TAG}_cfg2.err
timeout 900 python bench.py --workload cfg3 --steps 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "bench cfg3 rc=$?"; cut -c1-1500 gpurun_out/bench_${TAG}_cfg3.json; tail -3 gpurun_out/bench_${TAG}_cfg3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}_cfg2.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list_${TAG}_cfg2.log 2>&1; echo "ncu list cfg2 rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list_${TAG}_cfg3.log 2>&1; echo "ncu list cfg3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_project -c 1 -f -o gpurun_out/prof_project_${TAG}_cfg2 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full project rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -c 4 -f -o gpurun_out/prof_gemm_${TAG}_cfg2 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full2_${TAG}.log 2>&1; echo "ncu full gemm rc=$?"
timeout 120 tools/bin/umma_test > gpurun_out/umma_test_$TAG.log 2>&1; tail -1 gpurun_out/umma_test_$TAG.log
timeout 120 tools/bin/peaks_i8 > gpurun_out/peaks_i8_$TAG.json 2>&1
timeout 60 tools/bin/handoff_lat > gpurun_out/handoff_lat_$TAG.json 2>&1; cat gpurun_out/handoff_lat_$TAG.json
