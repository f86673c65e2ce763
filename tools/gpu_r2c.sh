#!/bin/bash
# Third GPU visit of round 2 (one GPU): suite with the nd = 0 two-block path and lean mode, sanitizer, traffic of the projection kernel.
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --workload cfg2 --steps 5 --no-cpu-baseline --no-fp64-extra > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --workload cfg3e --steps 3 --e2e-steps 2 --no-cpu-baseline --no-fp64-extra > gpurun_out/bench_${TAG}_cfg3e.json 2> gpurun_out/bench_${TAG}_cfg3e.err; echo "bench cfg3e rc=$?"
GEOBO_B200_LEAN_A=1 timeout 600 python bench.py --workload cfg3 --steps 2 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/bench_${TAG}_cfg3_lean.json 2> gpurun_out/bench_${TAG}_cfg3_lean.err; echo "bench cfg3 lean rc=$?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2c*.json")):
    try:
        d = json.load(open(p)); print(p, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "stages", {k: round(v, 2) for k, v in d["stage_ms"].items()}, "parity", (d.get("parity") or {}).get("max_err"), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"])
    except Exception as e:
        print(p, e, open(p.replace(".json", ".err")).read()[-600:])
PY
bash tools/gpu_sanitize.sh $TAG memcheck racecheck synccheck
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ozaki_project_kernel -c 1 --csv \
    --log-file gpurun_out/traffic_project_${TAG}_cfg3.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_traffic_${TAG}.log 2>&1; echo "ncu traffic rc=$?"; tail -3 gpurun_out/traffic_project_${TAG}_cfg3.csv
ls -la gpurun_out | tail -8
