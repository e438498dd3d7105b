#!/bin/bash
mkdir -p gpurun_out
for f in gpurun_tmp_libcfk_*.so; do
  echo "== $f =="
  CFK_LIBRARY=$PWD/$f timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
