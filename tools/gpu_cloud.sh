#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (warp form) =="
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_cloud.log
echo "== A/B =="
CFK_CLOUD_MODE=block timeout -k 10 600 python tools/ab_modes.py index_cap_mult=2 --steps 4 2>/dev/null | tee gpurun_out/ab_cloud.jsonl
timeout -k 10 600 python tools/ab_modes.py index_cap_mult=2,4,8 --steps 4 2>/dev/null | tee -a gpurun_out/ab_cloud.jsonl
