#!/bin/bash
# ncu evidence: launch lists (cfg2, cfg3), full captures of the projection kernel, the int8 GEMM kernels and the dense assembly kernel
TAG=${1:-p}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}_cfg2.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list_${TAG}_cfg2.log 2>&1; echo "ncu list cfg2 rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list_${TAG}_cfg3.log 2>&1; echo "ncu list cfg3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_project -c 1 -f -o gpurun_out/prof_project_${TAG}_cfg2 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full project rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:create_cov_grid_gather -s 3 -c 1 -f -o gpurun_out/prof_assembly_${TAG} \
    python tools/assembly_bench.py 32 32 16 > gpurun_out/ncu_asm_${TAG}.log 2>&1; echo "ncu full assembly rc=$?"
ls -la gpurun_out | tail -8
