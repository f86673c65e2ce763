#!/bin/bash
# First GPU visit of round 2 (one GPU, about 12 minutes): everything written in round 1 / session 3 without a GPU.
#   1. full GPU suite (new: align_drill, create_synsurvey, full-size parity fixtures cfg2 / cfg3, GEOBO_B200_CHOL_OUTER test,
#      structure: kron / compact -- tests/test_kron.py, tests/test_compact.py)
#   2. smoke, default bench line, cfg3 bench line
#   3. two-level Cholesky: cfg3 bench with GEOBO_B200_CHOL_OUTER=4 and a full ncu capture of its first gemm_f64 launches
#      (per outer block: panel solves, left-looking updates, then the K = 512 trailing update -- the launch with the largest grid)
# usage: tools/gpu_round2_first.sh <tag>;  2-GPU follow-up: gpurun --gpus 2 -- 'python -m pytest tests/test_multigpu.py tests/test_kron.py -m gpu -q'
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -60 gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/fullsize_parity.json 2>/dev/null
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_${TAG}_cfg2.json
timeout 900 python bench.py --workload cfg3 --steps 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "bench cfg3 rc=$?"; cut -c1-1200 gpurun_out/bench_${TAG}_cfg3.json
GEOBO_B200_CHOL_OUTER=4 timeout 900 python bench.py --workload cfg3 --steps 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3_outer4.json 2> gpurun_out/bench_${TAG}_cfg3_outer4.err; echo "bench cfg3 outer4 rc=$?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_*cfg3*.json")):
    try:
        d = json.load(open(p)); print(p, "chol ms", d["stage_ms"]["chol"], "value", d["value"])
    except Exception as e:
        print(p, e)
PY
GEOBO_B200_CHOL_OUTER=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -c 16 -f \
    -o gpurun_out/prof_chol_outer4_${TAG}_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline \
    > gpurun_out/ncu_chol_outer4_${TAG}.log 2>&1; echo "ncu chol outer4 rc=$?"
# 4. structure-exploiting paths written in session 4: bench lines and captures of their kernels
for PREC in fp64 int8x5; do
  timeout 600 python bench.py --structure kron --precision $PREC --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg2_kron_$PREC.json 2> gpurun_out/bench_${TAG}_cfg2_kron_$PREC.err
  echo "bench cfg2 kron $PREC rc=$?"; cut -c1-600 gpurun_out/bench_${TAG}_cfg2_kron_$PREC.json
done
timeout 600 python bench.py --workload cfg1 --structure compact --precision fp64 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg1_compact.json 2> gpurun_out/bench_${TAG}_cfg1_compact.err; echo "bench cfg1 compact rc=$?"
timeout 600 python bench.py --workload cfg1 --precision fp64 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg1_dense.json 2> gpurun_out/bench_${TAG}_cfg1_dense.err; echo "bench cfg1 dense rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kron_(y|zx)_kernel' -c 4 -f -o gpurun_out/prof_kron_${TAG}_cfg2 \
    python bench.py --structure kron --precision fp64 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_kron_${TAG}.log 2>&1; echo "ncu kron rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_cfg2_kron.csv \
    python bench.py --structure kron --precision int8x5 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1; echo "launch list kron rc=$?"
timeout 900 python bench.py --workload cfg3 --structure fft --precision int8x5 --steps 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3_fft.json 2> gpurun_out/bench_${TAG}_cfg3_fft.err; echo "bench cfg3 fft rc=$?"; cut -c1-600 gpurun_out/bench_${TAG}_cfg3_fft.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 60 -c 12 -f -o gpurun_out/prof_fft_${TAG}_cfg3 \
    python bench.py --workload cfg3 --structure fft --precision int8x5 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_fft_${TAG}.log 2>&1; echo "ncu fft rc=$?"
ls -la gpurun_out | tail -12
