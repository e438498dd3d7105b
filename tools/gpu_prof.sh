#!/bin/bash
# ncu: full-set capture of the two dominant kernels + launch list of one bench step.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:'pair_candidates_kernel|docfreq_kernel' -s 6 -c 2 \
   -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_full.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
ls -la gpurun_out
