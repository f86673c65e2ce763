#!/bin/bash
# 8-GPU visit: BASELINE config 5 through the DENSE projection with the streamed contraction (memory log on)
TAG=${1:-r2l}; N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--no-cpu-baseline --no-fp64-extra"
GEOBO_B200_MEMLOG=1 timeout 900 $TR --master-port 29631 bench.py --gpus $N --workload cfg5 --steps 1 --warmup 1 --min-warmup 1 --e2e-steps 1 --acq-sweep $B > gpurun_out/bench_${TAG}_cfg5_dense_n$N.json 2> gpurun_out/bench_${TAG}_cfg5_dense_n$N.err; echo "cfg5 dense n$N rc=$?"; grep "geobo_b200 rank [07]\]" gpurun_out/bench_${TAG}_cfg5_dense_n$N.err | tail -8; grep -i "error\|memory" gpurun_out/bench_${TAG}_cfg5_dense_n$N.err | head -5 | cut -c1-300
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2l*.json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1]); r = d["roofline"]
        print(p, "n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "stages", {k: round(v, 1) for k, v in d["stage_ms"].items() if v > 0.5}, "parity", (d.get("parity") or {}).get("max_err"),
              "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "ksteps", r.get("k_steps_visited_frac"), "bytes", d["impl_config"]["device_bytes"], d.get("checks"), d.get("acquisition_sweep"))
    except Exception as e:
        print(p, e)
PY
