#!/bin/bash
mkdir -p gpurun_out
LIB=${1:-}
if [ -n "$LIB" ]; then export CFK_LIBRARY=$PWD/$LIB; fi
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:'pair_candidates_kernel' -s 3 -c 1 \
   -f -o gpurun_out/prof_pc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pc.log 2>&1
tail -3 gpurun_out/ncu_pc.log
