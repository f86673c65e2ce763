#!/bin/bash
# usage: tools/gpu_mgpu.sh <tag> <ngpus> [workload]
TAG=$1; NG=$2; WL=${3:-cfg2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_mgpu_$TAG.log 2>&1; echo "pytest mgpu rc=$?"; tail -15 gpurun_out/pytest_mgpu_$TAG.log
for n in 1 $NG; do
  if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544"; fi
  NCCL_DEBUG=WARN timeout 1500 $L bench.py --gpus $n --workload $WL --steps 3 --e2e-steps 2 --no-cpu-baseline $EXTRA > gpurun_out/bench_${TAG}_${WL}_n$n.json 2> gpurun_out/bench_${TAG}_${WL}_n$n.err; echo "bench n=$n rc=$?"; tail -1 gpurun_out/bench_${TAG}_${WL}_n$n.json; tail -3 gpurun_out/bench_${TAG}_${WL}_n$n.err
done
