#!/bin/bash
# parity tests on the default build, then bench of the default + tuning variants (gpurun_tmp_libcfk_*.so)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== default =="
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_default.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['config']['edges'], d['config']['pair_candidates'])"
for f in gpurun_tmp_libcfk_*.so; do
  echo "== $f =="
  CFK_LIBRARY=$PWD/$f timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['config']['edges'], d['config']['pair_candidates'])"
done
