#!/bin/bash
# Final single-GPU visit: suite under -x, smoke, default bench, and the Cholesky trailing update with the grid / operand shapes of
# rank 3 of an 8-GPU block-cyclic factorisation (GEOBO_B200_CHOL_EMULATE, profiling aid) captured with ncu on one GPU
TAG=${1:-r2o}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 5 > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_default.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_r2o_default.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e", d["e2e"]["value"], d["e2e"]["first_call"], "parity", d["parity"]["max_err"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
PY
M="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
GEOBO_B200_CHOL_EMULATE=8,3 timeout 600 ncu --metrics $M --clock-control none -k regex:"gemm_f64_kernel<0" -c 170 --csv --log-file gpurun_out/chol_emulate8_${TAG}_cfg3.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_chol_emulate_${TAG}.log 2>&1; echo "ncu chol emulate list rc=$?"
GEOBO_B200_CHOL_EMULATE=8,3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_f64_kernel<0" --launch-skip 136 -c 6 -f -o gpurun_out/prof_chol_emulate8_${TAG}_cfg3 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_chol_emulate_full_${TAG}.log 2>&1; echo "ncu chol emulate full rc=$?"
ls -la gpurun_out | tail -5
