#!/bin/bash
# GPU box: ncu evidence of round 2. Usage: tools/profile_r2.sh   (writes gpurun_out/*_r2*)
mkdir -p gpurun_out
export PYTHONPATH=.
# 1. launch list of two steps of the bench workload, graphs off so that the solver stage shows launch by launch
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_pyramid.csv \
    python tools/profile_region.py pyramid 447 14 2 TGS_Soft 0 > gpurun_out/launches_r2_pyramid.log 2>&1
# 2. --set full of the dominant kernels of a step (persistent solver, narrow phase, pair queries, refit)
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'s2bPersistentSolveT|s2bUpdateContactsKernel|s2bFindPairs$|s2bRefit' -c 4 -f -o gpurun_out/step_r2 \
    python tools/profile_region.py pyramid 447 14 1 TGS_Soft 0 > gpurun_out/step_r2.log 2>&1
ncu -i gpurun_out/step_r2.ncu-rep --page raw --csv > gpurun_out/step_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/step_r2.ncu-rep --page details > gpurun_out/step_r2_details.txt 2>/dev/null
ncu -i gpurun_out/step_r2.ncu-rep --page source --csv -k regex:s2bPersistentSolveT > gpurun_out/step_r2_solve_source.csv 2>/dev/null
rm -f gpurun_out/step_r2.ncu-rep   # (gpurun_out travels back only below 64 MiB: keep the exported pages)
# 3. the launch-by-launch twin of the solver stage (S2B_PERSISTENT=0): integrate / gather body passes and colour passes as
#    kernels of their own
S2B_PERSISTENT=0 ncu --profile-from-start off --set full --clock-control none -k regex:'s2bBodyPassKernel|s2bRangePassKernel' -c 14 -f \
    -o gpurun_out/twin_r2 python tools/profile_region.py pyramid 447 14 1 TGS_Soft 0 > gpurun_out/twin_r2.log 2>&1
ncu -i gpurun_out/twin_r2.ncu-rep --page raw --csv > gpurun_out/twin_r2_raw.csv 2>/dev/null
rm -f gpurun_out/twin_r2.ncu-rep
# 4. per-colour kernel at scale, plain vs TMA bulk staging
python tools/color_kernel_probe.py 2600 > gpurun_out/colour_plain_r2.log 2>&1
S2B_COLOR_KERNEL=bulk python tools/color_kernel_probe.py 2600 > gpurun_out/colour_bulk_r2.log 2>&1
# 5. production step at 2 M boxes
python tools/production_at_scale.py 2000 8 > gpurun_out/production_2M_r2.log 2>&1
ls -la gpurun_out | tail -12
