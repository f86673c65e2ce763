#!/bin/bash
# multi-GPU sanity of the final code: usage tools/gpu_r2q.sh <ngpus> <tag>
N=${1:-2}; TAG=${2:-r2q}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py tests/test_kron.py -m gpu -q -k "two_rank" > gpurun_out/pytest_mgpu_${TAG}.log 2>&1; echo "pytest 2-rank rc=$?"; tail -3 gpurun_out/pytest_mgpu_${TAG}.log
fi
timeout 600 $TR --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_cfg3_n$N.json 2> gpurun_out/bench_${TAG}_cfg3_n$N.err; echo "cfg3 n$N rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_cfg3_n$N.err
timeout 300 $TR --master-port 29641 bench.py --impl reference --gpus $N --steps 2 --warmup 0 > gpurun_out/bench_${TAG}_ref_n$N.json 2> gpurun_out/bench_${TAG}_ref_n$N.err; echo "ref n$N rc=$?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/bench_r2q*_n[0-9].json")):
    try:
        d = json.loads([l for l in open(p) if l.startswith("{")][-1])
        print(p, "n", d["n_gpus"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "stages", {k: round(v, 1) for k, v in d.get("stage_ms", {}).items() if v > 0.5}, "parity", (d.get("parity") or {}).get("max_err"),
              "e2e", round(d["e2e"]["value"], 1), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(p, e)
PY
