#!/bin/bash
# GPU box: ncu evidence of round 2, final state (after the store / position folds, the Kempe pass and the hub rule).
# Usage: tools/profile_r2b.sh   (writes gpurun_out/*_r2b*)
mkdir -p gpurun_out
export PYTHONPATH=.
# 1. launch list of two steps of the bench workload, graphs off so that the solver stage shows launch by launch
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b_pyramid.csv \
    python tools/profile_region.py pyramid 447 14 2 TGS_Soft 0 > gpurun_out/launches_r2b_pyramid.log 2>&1
python tools/launch_list_summary.py gpurun_out/launches_r2b_pyramid.csv 60 > gpurun_out/launches_r2b_pyramid_summary.txt
# 2. --set full of the dominant kernels of a step (persistent solver, narrow phase, pair queries, refit)
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'s2bPersistentSolveT|s2bUpdateContactsKernel|s2bFindPairs$|s2bRefit' -c 4 -f -o gpurun_out/step_r2b \
    python tools/profile_region.py pyramid 447 14 1 TGS_Soft 0 > gpurun_out/step_r2b.log 2>&1
ncu -i gpurun_out/step_r2b.ncu-rep --page raw --csv > gpurun_out/step_r2b_raw.csv 2>/dev/null
ncu -i gpurun_out/step_r2b.ncu-rep --page details > gpurun_out/step_r2b_details.txt 2>/dev/null
rm -f gpurun_out/step_r2b.ncu-rep   # (gpurun_out travels back only below 64 MiB: keep the exported pages)
# 3. where the persistent kernel spends its time (stamps after every grid barrier), three scenes
python tools/trace_solve.py 447 pyramid 12 > gpurun_out/trace_pyramid_r2b.log 2>&1
python tools/trace_solve.py 256 field 12 > gpurun_out/trace_field_r2b.log 2>&1
python tools/trace_solve.py 100 tumbler 300 > gpurun_out/trace_tumbler_r2b.log 2>&1
# 4. launch list of two steps of the 10 k-box tumbler once the pile has formed
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b_tumbler.csv \
    python tools/profile_region.py tumbler 100 300 2 TGS_Soft 0 > gpurun_out/launches_r2b_tumbler.log 2>&1
python tools/launch_list_summary.py gpurun_out/launches_r2b_tumbler.csv 40 > gpurun_out/launches_r2b_tumbler_summary.txt
# 5. production step at 2 M boxes
python tools/production_at_scale.py 2000 8 > gpurun_out/production_2M_r2b.log 2>&1
ls -la gpurun_out | tail -14
