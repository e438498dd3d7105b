#!/bin/bash
# Stage A iteration: docfreq tests, then tools/stage_a_bench.py with the in-tree library and every variant.  -> gpurun_out/
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "docfreq" 2>&1 | tail -6 | tee gpurun_out/pytest_a.log
: > gpurun_out/stage_a.jsonl
for f in centroflye_b200/libcfk.so gpurun_tmp_libcfk_*.so; do
  [ -e "$f" ] || continue
  echo "== $f ==" | tee -a gpurun_out/stage_a.jsonl
  CFK_LIBRARY=$PWD/$f timeout -k 10 600 python tools/stage_a_bench.py ${ARGS:---check} 2> gpurun_out/stage_a.err | tee -a gpurun_out/stage_a.jsonl
  tail -3 gpurun_out/stage_a.err
done
