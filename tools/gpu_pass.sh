#!/bin/bash
# GPU pass: parity tests, smoke, small + full bench.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest gpu =="
timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
echo "== smoke =="
timeout -k 10 600 python __graft_entry__.py --smoke 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "== bench scale 0.1 =="
timeout -k 10 900 python bench.py --scale 0.1 --steps 3 --warmup 3 --cpu-seconds 3 > gpurun_out/bench_s01.json 2> gpurun_out/bench_s01.err
tail -5 gpurun_out/bench_s01.err; cat gpurun_out/bench_s01.json
echo "== bench full =="
timeout -k 10 1200 python bench.py --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
