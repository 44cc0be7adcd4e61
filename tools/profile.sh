#!/bin/bash
# GPU box: ncu evidence for bench.py's workload. Usage: tools/profile.sh <tag>
# 1. launch list of two timed steps (device time per launch; cold-cache and serialised: compare SHARES)
# 2. --set full capture of the persistent solver kernel (DRAM traffic, stalls, source page)
# 3. --set full capture of the per-colour kernel at scale (the roofline probe)
TAG=${1:-r1}
mkdir -p gpurun_out
export PYTHONPATH=.
export S2B_GRAPH=0   # ncu lists the kernels of a graph launch too, but keep the stage readable launch by launch
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 6 --no-cpu-baseline --no-colour-probe > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:s2bPersistentSolve -s 8 -c 1 -f -o gpurun_out/solve_${TAG} \
    python bench.py --steps 2 --warmup 6 --no-cpu-baseline --no-colour-probe > gpurun_out/solve_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:s2bTgsSoftColorKernel -s 4 -c 1 -f -o gpurun_out/colour_${TAG} \
    python tools/color_kernel_probe.py 2600 > gpurun_out/colour_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
