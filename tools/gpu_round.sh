#!/bin/bash
# Full single-GPU pass: parity tests, smoke, bench (both arms), ncu launch list + full capture.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest gpu =="
timeout -k 10 1200 python -m pytest tests -m gpu -x -q --durations=12 2>&1 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke =="
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench full =="
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== bench reference =="
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-seconds 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== ncu launches =="
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_full.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
echo "== ncu full =="
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'pair_sketch_kernel|docfreq_resident_kernel|cloud_build_warp_kernel|pair_join_kernel|table_select_kernel' -s 16 -c 5 \
   -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
