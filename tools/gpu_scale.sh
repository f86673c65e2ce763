#!/bin/bash
# usage: tools/gpu_scale.sh <tag> <ngpus>  -- scaling lines at N GPUs: cfg2 (default), cfg3, cfg4 (needs 8 GPUs)
TAG=$1; NG=$2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544"
run() { # workload steps e2e_steps timeout
  NCCL_DEBUG=WARN timeout $4 $L bench.py --gpus $NG --workload $1 --steps $2 --e2e-steps $3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$1_n$NG.json 2> gpurun_out/bench_${TAG}_$1_n$NG.err
  echo "bench $1 n=$NG rc=$?"; tail -1 gpurun_out/bench_${TAG}_$1_n$NG.json | cut -c1-1500; grep -v "nanvar\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_${TAG}_$1_n$NG.err | tail -4
}
run cfg2 5 3 600
run cfg3 3 1 600
if [ "$NG" -ge 8 ]; then run cfg4 1 1 900; fi
# opt-in structured projections, reported separately: STRUCT=1 tools/gpu_scale.sh <tag> <ngpus>
if [ -n "$STRUCT" ]; then
  runs() { # workload structure steps timeout
    NCCL_DEBUG=WARN timeout $4 $L bench.py --gpus $NG --workload $1 --structure $2 --steps $3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_$1_$2_n$NG.json 2> gpurun_out/bench_${TAG}_$1_$2_n$NG.err
    echo "bench $1 $2 n=$NG rc=$?"; tail -1 gpurun_out/bench_${TAG}_$1_$2_n$NG.json | cut -c1-800
  }
  GEOBO_B200_MGPU_STRUCTURED=1 timeout 600 $L tests/mgpu_check.py > gpurun_out/mgpu_structured_${TAG}_n$NG.log 2>&1; echo "mgpu structured rc=$?"; tail -2 gpurun_out/mgpu_structured_${TAG}_n$NG.log
  runs cfg2 kron 5 600
  runs cfg3 fft 3 600
  if [ "$NG" -ge 8 ]; then runs cfg4 kron 1 900; fi
fi
