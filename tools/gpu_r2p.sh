#!/bin/bash
# Cholesky trailing update with the grid / operand shapes of rank 3 of an 8-GPU block-cyclic factorisation, captured on one GPU
TAG=${1:-r2p}
mkdir -p gpurun_out
M="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
GEOBO_B200_CHOL_EMULATE=8,3 timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_f64_kernel -c 175 --csv --log-file gpurun_out/chol_emulate8_${TAG}_cfg3.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_chol_emulate_${TAG}.log 2>&1; echo "ncu chol emulate list rc=$?"
GEOBO_B200_CHOL_EMULATE=8,3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel --launch-skip 136 -c 5 -f -o gpurun_out/prof_chol_emulate8_${TAG}_cfg3 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_chol_emulate_full_${TAG}.log 2>&1; echo "ncu chol emulate full rc=$?"
wc -l gpurun_out/chol_emulate8_${TAG}_cfg3.csv; ls -la gpurun_out | tail -4
