#!/bin/bash
# ncu evidence for one bench step: launch list (gpu__time_duration) + full-set capture of the hot kernels.  Outputs -> gpurun_out/
mkdir -p gpurun_out
echo "== ncu launches =="
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_full.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
echo "== ncu full =="
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
   -k regex:'pair_sketch_kernel|docfreq_kernel|docfreq_count_kernel|docfreq_emit_kernel|cloud_build_kernel|pair_join_kernel|table_select_kernel' -s ${NCU_SKIP:-18} -c ${NCU_COUNT:-6} \
   -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
