#!/bin/bash
# 2-GPU visit: full suite (balanced shards, sharded lean regeneration, streamed contraction) + config 4 at 2 GPUs
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--no-cpu-baseline --no-fp64-extra"
timeout 900 $TR --master-port 29621 bench.py --gpus 2 --workload cfg4 --steps 2 --e2e-steps 1 $B > gpurun_out/bench_${TAG}_cfg4_n2.json 2> gpurun_out/bench_${TAG}_cfg4_n2.err; echo "cfg4 n2 rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_cfg4_n2.err
GEOBO_B200_STREAM_A8=1 GEOBO_B200_LEAN_A=1 timeout 600 python bench.py --workload cfg3e --steps 2 --e2e-steps 1 $B > gpurun_out/bench_${TAG}_cfg3e_streamed.json 2> gpurun_out/bench_${TAG}_cfg3e_streamed.err; echo "cfg3e streamed rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_cfg3e_streamed.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2j*.json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1]); r = d["roofline"]
        print(p, "n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "stages", {k: round(v, 1) for k, v in d["stage_ms"].items() if v > 0.5}, "parity", (d.get("parity") or {}).get("max_err"),
              "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "ksteps", r.get("k_steps_visited_frac"), "bytes", d["impl_config"]["device_bytes"])
    except Exception as e:
        print(p, e)
PY
