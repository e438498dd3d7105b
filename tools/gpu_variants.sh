#!/bin/bash
# Tuning: tools/ab_modes.py $SWITCH with the in-tree library and with every gpurun_tmp_libcfk_*.so variant
# (built by centroflye_b200.build.build_variant).  $TESTS: pytest -k expression run against every variant first.
mkdir -p gpurun_out; : > gpurun_out/variants.jsonl
for f in centroflye_b200/libcfk.so gpurun_tmp_libcfk_*.so; do
  [ -e "$f" ] || continue
  echo "== $f =="
  if [ -n "$TESTS" ]; then CFK_LIBRARY=$PWD/$f timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "$TESTS" 2>&1 | tail -2; fi
  CFK_LIBRARY=$PWD/$f timeout -k 10 600 python tools/ab_modes.py "${SWITCH:-docfreq_mode=resident}" --steps ${STEPS:-4} 2>/dev/null | sed "s|^{|{\"lib\": \"$f\", |" | tee -a gpurun_out/variants.jsonl
done
