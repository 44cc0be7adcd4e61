#!/bin/bash
# GPU box: ncu evidence for bench.py's workload. Usage: tools/profile.sh <tag>
# 1. launch list of two timed steps (device time per launch; cold-cache and serialised: compare SHARES)
# 2. --set full capture of the persistent solver kernel (DRAM traffic, stalls, source page)
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:s2bPersistentSolve -s 4 -c 1 -f -o gpurun_out/solve_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/solve_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
