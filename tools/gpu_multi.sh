#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): sharded-vs-oracle parity, then bench at N ranks.  Outputs -> gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== parity (world=$N) =="
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/multi_check_n$N.log | grep -i "error\|assert\|multi-gpu ok\|File" | head -20
echo "== bench (world=$N) =="
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err | cut -c1-300; cat gpurun_out/bench_n$N.json | cut -c1-3000
