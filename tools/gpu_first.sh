#!/bin/bash
# First GPU pass: parity tests, smoke, small + full bench, launch list.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest gpu ==" 
timeout -k 10 1500 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/pytest_gpu.log; tail -60 gpurun_out/pytest_gpu.log
echo "== smoke =="
timeout -k 10 600 python __graft_entry__.py --smoke 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "== bench scale 0.1 =="
timeout -k 10 900 python bench.py --scale 0.1 --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_s01.json 2> gpurun_out/bench_s01.err
tail -5 gpurun_out/bench_s01.err; cat gpurun_out/bench_s01.json
echo "== bench full =="
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== ncu launch list (scale 0.1) =="
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s01.csv python bench.py --scale 0.1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "== sanitizer (edge cases) =="
timeout -k 10 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "edge_cases" 2>&1 | tail -15 | tee gpurun_out/sanitizer.log
