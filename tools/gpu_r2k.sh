#!/bin/bash
# 8-GPU visit: balanced shards (cfg3, cfg4) and BASELINE config 5 through the DENSE projection with the streamed contraction
TAG=${1:-r2k}; N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--no-cpu-baseline --no-fp64-extra"
timeout 400 $TR --master-port 29611 bench.py --gpus $N --steps 5 --e2e-steps 3 $B > gpurun_out/bench_${TAG}_cfg3_n$N.json 2> gpurun_out/bench_${TAG}_cfg3_n$N.err; echo "cfg3 n$N rc=$?"; tail -c 200 gpurun_out/bench_${TAG}_cfg3_n$N.err
timeout 600 $TR --master-port 29621 bench.py --gpus $N --workload cfg4 --steps 2 --e2e-steps 1 $B > gpurun_out/bench_${TAG}_cfg4_n$N.json 2> gpurun_out/bench_${TAG}_cfg4_n$N.err; echo "cfg4 n$N rc=$?"; tail -c 200 gpurun_out/bench_${TAG}_cfg4_n$N.err
timeout 900 $TR --master-port 29631 bench.py --gpus $N --workload cfg5 --steps 1 --e2e-steps 1 --acq-sweep $B > gpurun_out/bench_${TAG}_cfg5_dense_n$N.json 2> gpurun_out/bench_${TAG}_cfg5_dense_n$N.err; echo "cfg5 dense n$N rc=$?"; tail -c 800 gpurun_out/bench_${TAG}_cfg5_dense_n$N.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2k*.json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1]); r = d["roofline"]
        print(p, "n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "stages", {k: round(v, 1) for k, v in d["stage_ms"].items() if v > 0.5}, "parity", (d.get("parity") or {}).get("max_err"),
              "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "ksteps", r.get("k_steps_visited_frac"), "bytes", d["impl_config"]["device_bytes"], d.get("checks"), d.get("acquisition_sweep"))
    except Exception as e:
        print(p, e)
PY
