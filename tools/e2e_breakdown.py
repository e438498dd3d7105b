#!/usr/bin/env python
"""Where does the end-to-end step (pinned host buffers -> host results) spend its time?  Host timers with a
synchronize after every phase, plus raw pinned H2D / D2H bandwidth of the box.  Diagnostic only."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from centroflye_b200.engine import Engine
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    eng = Engine("cuda:0")
    unit, batch, units = bench.make_inputs(scale)
    k = bench.PARAMS["k"]
    lo, hi = bench.band()
    P = bench.PARAMS

    def sync():
        torch.cuda.synchronize()

    # raw copy bandwidth
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); sync()
        t = time.perf_counter()
        for _ in range(4):
            fn()
        sync()
        print(f"pinned {name}: {4 * n / (time.perf_counter() - t) / 1e9:.1f} GB/s")

    for it in range(4):
        T = {}
        sync()
        t0 = time.perf_counter()
        reads = eng.upload_reads(batch, k); sync(); T["upload_reads"] = time.perf_counter()
        dunits = eng.upload_units(units, k); sync(); T["upload_units"] = time.perf_counter()
        table = eng.count_docfreq(reads, k); sync(); T["docfreq"] = time.perf_counter()
        rare = eng.table_select(table, lo, hi, P["max_nonuniq"]); sync(); T["select"] = time.perf_counter()
        del table
        index = eng.build_index(rare); sync(); T["index"] = time.perf_counter()
        csr = eng.build_clouds(reads, dunits, k, index); sync(); T["clouds"] = time.perf_counter()
        res = eng.dist_edges(csr, dunits.unit_last, index.n, P["min_d"], P["max_d"], P["min_coverage"]); sync()
        T["dist_edges"] = time.perf_counter()
        out = eng.to_host(selected=res.selected, edges=res.edges, unit_ptr=csr.unit_ptr, ids=csr.ids,
                          rare_keys=index.sorted_keys)
        sync(); T["to_host"] = time.perf_counter()
        prev = t0
        parts = []
        for name, t in T.items():
            parts.append(f"{name} {1e3 * (t - prev):.2f}")
            prev = t
        print(f"iter {it}: total {1e3 * (prev - t0):.2f} ms | " + " | ".join(parts))
        print("   d2h bytes", sum(x.numel() * x.element_size() for x in out.values()))


if __name__ == "__main__":
    main()
