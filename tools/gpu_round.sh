#!/bin/bash
# Full single-GPU measurement pass: bench (all configs, both arms), ncu launch list + full capture.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== bench cenx =="
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -2 gpurun_out/bench_full.err
echo "== bench stream =="
timeout -k 10 600 python bench.py --config stream --steps 8 --warmup 3 > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
tail -2 gpurun_out/bench_stream.err
echo "== bench cen6 =="
timeout -k 10 600 python bench.py --config cen6 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cen6.json 2> gpurun_out/bench_cen6.err
tail -2 gpurun_out/bench_cen6.err
echo "== bench reference =="
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -2 gpurun_out/bench_ref.err
echo "== ncu launches =="
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_full.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
echo "== ncu full =="
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'pair_sketch_kernel|docfreq_emit_kernel|docfreq_count_kernel|cloud_build_warp_kernel|pair_join_kernel' -s 15 -c 5 \
   -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | tail -12
