#!/bin/bash
# ncu --set full capture of the stage-A kernels (one launch each, warm) + a launch list of one step.  -> gpurun_out/
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'docfreq_emit_kernel|docfreq_apply_kernel' -s 6 -c 2 \
   -f -o gpurun_out/prof_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_a.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_a.log 2>&1
tail -2 gpurun_out/ncu_launches_a.log
