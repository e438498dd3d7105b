#!/bin/bash
# Quick pass: kernel + parity tests, short bench, optional racecheck of the stage-C tests.  Outputs -> gpurun_out/
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout -k 10 1200 python -m pytest tests -m gpu -x -q --durations=12 2>&1 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
echo "== bench =="
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/bench_quick.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read())
print(d['ms_per_step'], d['stage_ms'], d['config']['edges'], d['config']['pair_candidates'], d['e2e']['ms_per_step'])
PY
for f in gpurun_tmp_libcfk_*.so; do
  [ -e "$f" ] || continue
  echo "== $f =="
  CFK_LIBRARY=$PWD/$f timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['config']['edges'], d['config']['pair_candidates'])"
done
if [ -n "$RACECHECK" ]; then
echo "== racecheck =="
timeout -k 10 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sketch_random or table_splitting" 2>&1 | tail -12 | tee gpurun_out/racecheck.log
fi
