#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
for WL in cfg2 cfg3; do
timeout 900 python bench.py --workload $WL --steps 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${WL}.json 2> gpurun_out/bench_${TAG}_${WL}.err; echo "bench $WL rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_${WL}.json").read().strip().splitlines()[-1])
print("$WL", round(d["value"]), "vox/s", {k: round(v,2) for k,v in d["stage_ms"].items()}, d["clocks"], "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), "launches", d["gpu_launches"])
PY
tail -2 gpurun_out/bench_${TAG}_${WL}.err
done
