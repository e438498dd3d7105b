#!/usr/bin/env python
"""A/B of engine switches on the bench workload with ONE input generation:
    python tools/ab_modes.py docfreq_mode=tiled,resident [--steps 5] [--scale 1.0]
Prints ms per step and per stage for every value of the switch (CUDA events, L2 flushed between steps)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("switch")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    from centroflye_b200.engine import Engine
    name, values = args.switch.split("=")
    eng = Engine("cuda:0")
    unit, batch, units = bench.make_inputs(args.scale)
    P = bench.PARAMS
    lo, hi = bench.band()
    reads, dunits = eng.upload_reads(batch, P["k"]), eng.upload_units(units, P["k"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    for v in values.split(","):
        setattr(eng, name, type(getattr(eng, name))(v))
        step, stages, sig = [], {}, None
        for i in range(args.warmup + args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            eng.events = []
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            index, csr, res = eng.recruit(reads, dunits, P["k"], lo, hi, P["max_nonuniq"], P["min_d"], P["max_d"],
                                          P["min_coverage"])
            b.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                step.append(a.elapsed_time(b))
                for s, ms in eng.stage_times_ms().items():
                    stages.setdefault(s, []).append(ms)
            eng.events = None
            sig = (index.n, csr.n_entries, int(res.edges.shape[0]), int(res.selected.numel()), res.n_increments)
        print(json.dumps({"switch": name, "value": v, "ms_per_step": float(np.mean(step)),
                          "stage_ms": {s: round(float(np.mean(x)), 4) for s, x in stages.items()},
                          "signature": sig}), flush=True)


if __name__ == "__main__":
    main()
