#!/bin/bash
# Final single-GPU validation of round 2: suite as the driver runs it, smoke, both bench arms, launch list + full capture of the projection kernel
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref arm rc=$?"
timeout 900 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_default.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2m*.json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1])
        print(p, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1), "parity", (d.get("parity") or {}).get("max_err"), (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(p, e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}_cfg3.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > /dev/null 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_project_kernel -c 1 -f -o gpurun_out/prof_project_${TAG}_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_project_${TAG}.log 2>&1; echo "ncu project cfg2 rc=$?"
ls -la gpurun_out | tail -6
