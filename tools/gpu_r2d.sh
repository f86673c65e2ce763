#!/bin/bash
# Fourth GPU visit (one GPU): zero-digit culling + tile-round pacing of the projection kernel -- suite, A/B bench lines, traffic.
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
B="--e2e-steps 1 --no-cpu-baseline --no-fp64-extra"
timeout 300 python bench.py --steps 3 $B > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "cfg3 (pacing on) rc=$?"
GEOBO_B200_TILE_SYNC=0 timeout 300 python bench.py --steps 3 $B > gpurun_out/bench_${TAG}_cfg3_nosync.json 2> gpurun_out/bench_${TAG}_cfg3_nosync.err; echo "cfg3 (pacing off) rc=$?"
timeout 300 python bench.py --workload cfg3e --steps 3 $B > gpurun_out/bench_${TAG}_cfg3e.json 2> gpurun_out/bench_${TAG}_cfg3e.err; echo "cfg3e rc=$?"
GEOBO_B200_CULL=0 timeout 300 python bench.py --workload cfg3e --steps 3 $B > gpurun_out/bench_${TAG}_cfg3e_nocull.json 2> gpurun_out/bench_${TAG}_cfg3e_nocull.err; echo "cfg3e nocull rc=$?"
timeout 300 python bench.py --workload cfg2 --steps 5 $B > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload cfg1 --steps 5 $B > gpurun_out/bench_${TAG}_cfg1.json 2> gpurun_out/bench_${TAG}_cfg1.err; echo "cfg1 rc=$?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2d*.json")):
    try:
        d = json.load(open(p)); r = d["roofline"]
        print(p, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "project", round(d["stage_ms"]["project"], 2), "trsm", round(d["stage_ms"]["trsm"], 2), "parity", (d.get("parity") or {}).get("max_err"),
              "frac", round(r["frac"], 3), "ksteps", r.get("k_steps_visited_frac"), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"))
    except Exception as e:
        print(p, e, open(p.replace(".json", ".err")).read()[-600:])
PY
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ozaki_project_kernel -c 1 --csv \
    --log-file gpurun_out/traffic_project_${TAG}_cfg3.csv python bench.py --steps 1 --warmup 3 $B > gpurun_out/ncu_traffic_${TAG}.log 2>&1; echo "ncu traffic rc=$?"; tail -3 gpurun_out/traffic_project_${TAG}_cfg3.csv | cut -c100-400
timeout 600 bash tools/gpu_sanitize.sh $TAG memcheck
