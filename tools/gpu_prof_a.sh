#!/bin/bash
# ncu --set full capture of the stage-A kernels (one warm launch each) through tools/stage_a_bench.py.  -> gpurun_out/
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-docfreq_emit_kernel|docfreq_count_kernel}" -s ${SKIP:-4} -c ${COUNT:-2} \
   -f -o gpurun_out/prof_a python tools/stage_a_bench.py --modes ${MODES:-stream} --steps 1 --warmup 2 ${ARGS} > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/ncu_a.log
