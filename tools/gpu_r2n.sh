#!/bin/bash
# A-B of the alternating sorted tile order (one GPU, ~4 min)
TAG=${1:-r2n}
mkdir -p gpurun_out
B="--e2e-steps 1 --no-cpu-baseline --no-fp64-extra --steps 3"
timeout 200 python -m pytest tests/test_gpu_int8.py -m gpu -q -k "culling or lean or chunks" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
run() { NAME=$1; shift; env "$@" timeout 300 python bench.py $WL $B > gpurun_out/bench_${TAG}_$NAME.json 2> gpurun_out/bench_${TAG}_$NAME.err; echo "$NAME rc=$?"; }
WL=""
run cfg3_sync1 GEOBO_B200_TILE_SYNC=1
run cfg3_sync2 GEOBO_B200_TILE_SYNC=2
WL="--workload cfg3e"
run cfg3e_sync1 GEOBO_B200_TILE_SYNC=1
run cfg3e_sync2 GEOBO_B200_TILE_SYNC=2
WL="--workload cfg2"
run cfg2_sync1 GEOBO_B200_TILE_SYNC=1
run cfg2_sync2 GEOBO_B200_TILE_SYNC=2
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2n*.json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1]); r = d["roofline"]
        print(p, "value", round(d["value"]), "project", round(d["stage_ms"]["project"], 2), "frac", round(r["frac"], 3), "ksteps", round(r.get("k_steps_visited_frac"), 3), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"))
    except Exception as e:
        print(p, e, open(p.replace(".json", ".err")).read()[-600:])
PY
