#!/bin/bash
# Multi-GPU visit of round 2: usage tools/gpu_mgpu_r2.sh <ngpus> <tag>   (run under `gpurun --gpus N`)
N=${1:-2}; TAG=${2:-r2g}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py tests/test_kron.py -m gpu -q -k "two_rank" > gpurun_out/pytest_mgpu_${TAG}.log 2>&1; echo "pytest 2-rank rc=$?"; tail -4 gpurun_out/pytest_mgpu_${TAG}.log
fi
if [ "$N" = "4" ]; then
  timeout 300 python -m pytest tests/test_gpu_int8.py -m gpu -q -k "lean or chunks or culling" > gpurun_out/pytest_lean_${TAG}.log 2>&1; echo "pytest lean rc=$?"; tail -4 gpurun_out/pytest_lean_${TAG}.log
fi
B="--no-cpu-baseline --no-fp64-extra"
timeout 600 $TR --master-port 29611 bench.py --gpus $N --steps 5 --e2e-steps 3 $B > gpurun_out/bench_${TAG}_cfg3_n$N.json 2> gpurun_out/bench_${TAG}_cfg3_n$N.err; echo "cfg3 n$N rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_cfg3_n$N.err
timeout 1500 $TR --master-port 29621 bench.py --gpus $N --workload cfg4 --steps 2 --e2e-steps 1 $B > gpurun_out/bench_${TAG}_cfg4_n$N.json 2> gpurun_out/bench_${TAG}_cfg4_n$N.err; echo "cfg4 n$N rc=$?"; tail -c 600 gpurun_out/bench_${TAG}_cfg4_n$N.err
if [ "$N" = "8" ]; then
  timeout 1500 $TR --master-port 29631 bench.py --gpus $N --workload cfg5 --structure kron --steps 2 --e2e-steps 1 --acq-sweep $B > gpurun_out/bench_${TAG}_cfg5_kron_n$N.json 2> gpurun_out/bench_${TAG}_cfg5_kron_n$N.err; echo "cfg5 kron n$N rc=$?"; tail -c 800 gpurun_out/bench_${TAG}_cfg5_kron_n$N.err
fi
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_*_n[0-9].json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1]); r = d["roofline"]
        print(p, "n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "stages", {k: round(v, 1) for k, v in d["stage_ms"].items() if v > 0.5}, "parity", (d.get("parity") or {}).get("max_err"),
              "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "ksteps", r.get("k_steps_visited_frac"), "bytes", d["impl_config"]["device_bytes"])
    except Exception as e:
        print(p, e)
PY
nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv > gpurun_out/mem_${TAG}_n$N.txt 2>&1
