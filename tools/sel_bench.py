#!/usr/bin/env python
"""Time cfk_table_select alone on the bench workload's stage-A table: the rare band and a denser nomination band."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from centroflye_b200.engine import Engine

eng = Engine("cuda:0")
unit, batch, units = bench.make_inputs(1.0)
reads = eng.upload_reads(batch, 19)
table = eng.count_docfreq(reads, 19)
lo, hi = bench.band()
for name, (a, b, c) in {"rare": (lo, hi, 3), "ge5": (5, 0xFFFFFFFF, 0xFFFFFFFF), "ge2": (2, 0xFFFFFFFF, 0xFFFFFFFF)}.items():
    ts = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        keys = eng.table_select(table, a, b, c)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(json.dumps({"lib": os.environ.get("CFK_LIBRARY", "default"), "band": name, "n": int(keys.numel()), "ms": round(min(ts[2:]), 4)}))
