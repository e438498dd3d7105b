#!/bin/bash
# Multi-GPU bench at N ranks (run with gpurun --gpus N): configs[1] weak scaling with the parity leg, then the stream config.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps ${STEPS:-4} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -6 | cut -c1-200
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --config stream --steps 6 --warmup 3 > gpurun_out/bench_stream_n$N.json 2> gpurun_out/bench_stream_n$N.err
grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_stream_n$N.err | tail -3 | cut -c1-200
python - <<PY
import json
for f in ("gpurun_out/bench_n$N.json", "gpurun_out/bench_stream_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 2), "Gb/s", round(d["value"] / 1e9, 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "parity", d.get("parity_ok"))
        print("  ", {k: round(v, 2) for k, v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "no json:", e)
PY
