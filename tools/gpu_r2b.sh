#!/bin/bash
# Second GPU visit of round 2 (one GPU): suite, the driver-style default bench line (cfg3 + in-run parity + fp64 extra), both CPU arms,
# compute-sanitizer, Cholesky look-ahead timing, ncu capture of the projection kernel at 64x64x32 (traffic) and of stencil_kernel.
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref_cfg3.json 2> gpurun_out/bench_${TAG}_ref_cfg3.err; echo "ref arm rc=$?"; cut -c1-700 gpurun_out/bench_${TAG}_ref_cfg3.json
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "bench default rc=$?"; cut -c1-600 gpurun_out/bench_${TAG}_cfg3.json; tail -3 gpurun_out/bench_${TAG}_cfg3.err
GEOBO_B200_CHOL_LOOKAHEAD=1 timeout 900 python bench.py --steps 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/bench_${TAG}_cfg3_la.json 2> gpurun_out/bench_${TAG}_cfg3_la.err; echo "bench lookahead rc=$?"
timeout 900 python bench.py --workload cfg3e --steps 3 --e2e-steps 2 --no-cpu-baseline --no-fp64-extra > gpurun_out/bench_${TAG}_cfg3e.json 2> gpurun_out/bench_${TAG}_cfg3e.err; echo "bench cfg3e rc=$?"
timeout 900 python bench.py --workload cfg3e --structure kron --steps 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3e_kron.json 2> gpurun_out/bench_${TAG}_cfg3e_kron.err; echo "bench cfg3e kron rc=$?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2b*cfg3*.json")):
    try:
        d = json.load(open(p)); print(p, "value", round(d["value"]), "chol ms", d.get("stage_ms", {}).get("chol"), "parity", (d.get("parity") or {}).get("max_err"), "e2e", d["e2e"]["value"])
    except Exception as e:
        print(p, e)
PY
bash tools/gpu_sanitize.sh $TAG memcheck racecheck synccheck
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_project_kernel -c 1 -f -o gpurun_out/prof_project_${TAG}_cfg3 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_project_${TAG}.log 2>&1; echo "ncu project cfg3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stencil_kernel -c 2 -f -o gpurun_out/prof_stencil_${TAG}_cfg1 \
    python bench.py --workload cfg1 --structure compact --precision fp64 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_stencil_${TAG}.log 2>&1; echo "ncu stencil rc=$?"
ls -la gpurun_out | tail -8
