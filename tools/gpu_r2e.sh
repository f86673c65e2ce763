#!/bin/bash
# pacing slack / producer K mapping A-B on the 64x64x32 cubes (one GPU, ~3 min)
TAG=${1:-r2e}
mkdir -p gpurun_out
B="--e2e-steps 1 --no-cpu-baseline --no-fp64-extra --steps 3"
timeout 120 python -m pytest tests/test_gpu_int8.py -m gpu -q -k "culling or chunks or lean" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for MODE in 1 2 3; do
  GEOBO_B200_TILE_SYNC=$MODE timeout 300 python bench.py $B > gpurun_out/bench_${TAG}_cfg3_sync$MODE.json 2> gpurun_out/bench_${TAG}_cfg3_sync$MODE.err; echo "cfg3 sync$MODE rc=$?"
done
GEOBO_B200_CULL=0 timeout 300 python bench.py $B > gpurun_out/bench_${TAG}_cfg3_nocull.json 2> gpurun_out/bench_${TAG}_cfg3_nocull.err; echo "cfg3 nocull rc=$?"
for MODE in 1 2; do
  GEOBO_B200_TILE_SYNC=$MODE timeout 300 python bench.py --workload cfg3e $B > gpurun_out/bench_${TAG}_cfg3e_sync$MODE.json 2> gpurun_out/bench_${TAG}_cfg3e_sync$MODE.err; echo "cfg3e sync$MODE rc=$?"
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2e*.json")):
    try:
        d = json.load(open(p)); r = d["roofline"]
        print(p, "value", round(d["value"]), "project", round(d["stage_ms"]["project"], 2), "frac", round(r["frac"], 3), "ksteps", round(r.get("k_steps_visited_frac"), 3), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"))
    except Exception as e:
        print(p, e, open(p.replace(".json", ".err")).read()[-600:])
PY
GEOBO_B200_TILE_SYNC=2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ozaki_project_kernel -c 1 --csv \
    --log-file gpurun_out/traffic_project_${TAG}_cfg3_sync2.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_traffic_${TAG}.log 2>&1; echo "ncu traffic rc=$?"; tail -3 gpurun_out/traffic_project_${TAG}_cfg3_sync2.csv | cut -c100-400
