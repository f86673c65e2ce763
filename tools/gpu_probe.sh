#!/bin/bash
# First-contact probe of the GPU box: host/GPU facts, fp64 peaks, smoke, GPU parity tests.
mkdir -p gpurun_out
{
  echo "== host"; nproc; free -g | head -2; python -c "import os; print('cpu_count', os.cpu_count())"
  echo "== gpu"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  nvidia-smi topo -m 2>/dev/null | head -12
} > gpurun_out/probe.txt 2>&1
timeout 120 tools/bin/peaks > gpurun_out/peaks_fp64.json 2>&1
cat gpurun_out/peaks_fp64.json
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
