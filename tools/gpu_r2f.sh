#!/bin/bash
# sorted tile order A-B (one GPU, ~3 min)
TAG=${1:-r2f}
mkdir -p gpurun_out
B="--e2e-steps 1 --no-cpu-baseline --no-fp64-extra --steps 3"
timeout 200 python -m pytest tests/test_gpu_int8.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
run() { NAME=$1; shift; env "$@" timeout 300 python bench.py $WL $B > gpurun_out/bench_${TAG}_$NAME.json 2> gpurun_out/bench_${TAG}_$NAME.err; echo "$NAME rc=$?"; }
WL=""
run cfg3_sort_sync1 GEOBO_B200_TILE_SYNC=1
run cfg3_sort_sync2 GEOBO_B200_TILE_SYNC=2
run cfg3_nosort_sync2 GEOBO_B200_TILE_SYNC=2 GEOBO_B200_TILE_SORT=0
WL="--workload cfg3e"
run cfg3e_sort_sync1 GEOBO_B200_TILE_SYNC=1
run cfg3e_sort_sync2 GEOBO_B200_TILE_SYNC=2
WL="--workload cfg2"
run cfg2_sort_sync1 GEOBO_B200_TILE_SYNC=1
run cfg2_sort_sync2 GEOBO_B200_TILE_SYNC=2
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2f*.json")):
    try:
        d = json.load(open(p)); r = d["roofline"]
        print(p, "value", round(d["value"]), "project", round(d["stage_ms"]["project"], 2), "frac", round(r["frac"], 3), "ksteps", round(r.get("k_steps_visited_frac"), 3), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"))
    except Exception as e:
        print(p, e, open(p.replace(".json", ".err")).read()[-600:])
PY
GEOBO_B200_TILE_SYNC=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ozaki_project_kernel -c 1 --csv \
    --log-file gpurun_out/traffic_project_${TAG}_cfg3_sort_sync1.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-fp64-extra > gpurun_out/ncu_traffic_${TAG}.log 2>&1; echo "ncu traffic rc=$?"; tail -3 gpurun_out/traffic_project_${TAG}_cfg3_sort_sync1.csv | cut -c100-400
