#!/bin/bash
# int8 peak microbenchmark, bench lines for the int8 paths, ncu launch list + full capture of the tcgen05 projection kernel
TAG=${1:-oz}
mkdir -p gpurun_out
timeout 120 tools/bin/peaks_i8 > gpurun_out/peaks_i8_$TAG.json 2>&1; echo "peaks_i8 rc=$?"; cat gpurun_out/peaks_i8_$TAG.json
for PR in int8x6 int8x5; do
timeout 600 python bench.py --workload cfg2 --precision $PR --steps 5 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg2_$PR.json 2> gpurun_out/bench_${TAG}_cfg2_$PR.err; echo "bench $PR rc=$?"; cat gpurun_out/bench_${TAG}_cfg2_$PR.json; tail -3 gpurun_out/bench_${TAG}_cfg2_$PR.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_project -c 1 -f -o gpurun_out/prof_ozaki_${TAG}_cfg2 \
    python bench.py --workload cfg2 --precision int8x6 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 900 python tools/ozaki_check.py cfg3 > gpurun_out/ozaki_cfg3_$TAG.log 2>&1; echo "ozaki cfg3 rc=$?"; tail -4 gpurun_out/ozaki_cfg3_$TAG.log
ls -la gpurun_out | tail -8
