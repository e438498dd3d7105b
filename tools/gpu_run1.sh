#!/bin/bash
# Stage-A resident kernel: parity first, then A/B against the tiled kernel, then an ncu full capture of it.  Outputs -> gpurun_out/
mkdir -p gpurun_out
echo "== pytest stage A =="
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "docfreq or table_select" 2>&1 | tail -8 | tee gpurun_out/pytest_docfreq.log
echo "== pytest gpu (all) =="
timeout -k 10 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; tail -14 gpurun_out/pytest_gpu.log
echo "== A/B =="
timeout -k 10 600 python tools/ab_modes.py docfreq_mode=tiled,resident --steps 5 2> gpurun_out/ab.err | tee gpurun_out/ab_docfreq.jsonl
tail -2 gpurun_out/ab.err
echo "== ncu full (stage A) =="
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'docfreq_resident_kernel' -s 2 -c 1 \
   -f -o gpurun_out/prof_docfreq python tools/ab_modes.py docfreq_mode=resident --steps 1 --warmup 2 > gpurun_out/ncu_docfreq.log 2>&1
tail -2 gpurun_out/ncu_docfreq.log
ls -la gpurun_out
