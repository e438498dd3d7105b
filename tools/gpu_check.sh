#!/bin/bash
# Full GPU test suite, then tools/ab_modes.py $SWITCH (A/B of an engine switch on one input generation).  Outputs -> gpurun_out/
mkdir -p gpurun_out
echo "== pytest =="
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_occ.log
echo "== A/B =="
timeout -k 10 600 python tools/ab_modes.py ${SWITCH:-use_occ_last=0,1} --steps 4 2>/dev/null | tee gpurun_out/ab_occ.jsonl
