#!/bin/bash
TAG=${1:-oz}; PR=${2:-int8x5}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_int8.py -m gpu -x -q > gpurun_out/pytest_int8_$TAG.log 2>&1; echo "pytest int8 rc=$?"; tail -6 gpurun_out/pytest_int8_$TAG.log
for WL in cfg2 cfg3; do
timeout 900 python bench.py --workload $WL --precision $PR --steps 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_${WL}_$PR.json 2> gpurun_out/bench_${TAG}_${WL}_$PR.err; echo "bench $WL rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_${WL}_$PR.json").read().strip().splitlines()[-1])
print("$WL", round(d["value"]), "vox/s", {k: round(v,2) for k,v in d["stage_ms"].items()}, d["clocks"], "e2e", round(d["e2e"]["value"]))
PY
tail -2 gpurun_out/bench_${TAG}_${WL}_$PR.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}_cfg2_$PR.csv \
    python bench.py --workload cfg2 --precision $PR --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
