#!/bin/bash
# Quick A/B: targeted tests ($TESTS = pytest -k expression), then tools/ab_modes.py $SWITCH, optional ncu full capture of $NCU_KERNEL.  Outputs -> gpurun_out/
mkdir -p gpurun_out
echo "== pytest =="
timeout -k 10 900 python -m pytest tests -m gpu -x -q -k "${TESTS:-docfreq or table_select}" 2>&1 | tail -8 | tee gpurun_out/pytest_ab.log
echo "== A/B =="
timeout -k 10 600 python tools/ab_modes.py "${SWITCH:-docfreq_mode=tiled,resident}" --steps 5 2> gpurun_out/ab.err | tee gpurun_out/ab.jsonl
tail -2 gpurun_out/ab.err
if [ -n "$NCU_KERNEL" ]; then
echo "== ncu full =="
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU_KERNEL" -s 2 -c 1 \
   -f -o gpurun_out/prof_ab python tools/ab_modes.py "${NCU_SWITCH:-docfreq_mode=resident}" --steps 1 --warmup 2 > gpurun_out/ncu_ab.log 2>&1
tail -2 gpurun_out/ncu_ab.log
fi
