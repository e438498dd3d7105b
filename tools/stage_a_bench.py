#!/usr/bin/env python
"""Stage A alone on the bench workload: the two-phase kernels (emit + count) against the single-kernel form.
    python tools/stage_a_bench.py [--steps 5] [--scale 1.0] [--check] [--table]
Prints one JSON line per variant: ms of every kernel (CUDA events, L2 flushed between steps), Gbases/s, and with
--check whether rare set and full table equal those of the resident kernel."""
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--table", action="store_true", help="also write the dense table in phase 2")
    ap.add_argument("--modes", default="stream,resident")
    args = ap.parse_args()
    import torch
    from centroflye_b200.engine import Engine
    eng = Engine("cuda:0")
    unit, batch, units = bench.make_inputs(args.scale)
    P = bench.PARAMS
    k = P["k"]
    lo, hi = bench.band()
    reads = eng.upload_reads(batch, k)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    n_k = int(batch.n_bases - batch.n_reads * (k - 1))
    results = {}
    for mode in args.modes.split(","):
        eng.docfreq_mode = mode
        stages, total = {}, []
        for i in range(args.warmup + args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            eng.events = []
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if mode == "stream":
                out = eng.docfreq_stream(reads, k, band=(lo, hi, P["max_nonuniq"]), want_table=args.table)
                if out is None:
                    raise SystemExit("stream mode fell back")
                rare = out[0]
            else:
                rare = eng.rare_kmers(reads, k, lo, hi, P["max_nonuniq"])
            b.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                total.append(a.elapsed_time(b))
                for s, ms in eng.stage_times_ms().items():
                    stages.setdefault(s, []).append(ms)
            eng.events = None
        st = {s: round(float(np.mean(x)), 4) for s, x in stages.items()}
        kern = sum(st.values())
        results[mode] = np.sort(rare.cpu().numpy().view(np.uint64))
        print(json.dumps({"mode": mode, "call_ms": round(float(np.mean(total)), 3), "kernel_ms": round(kern, 3),
                          "stage_ms": st, "gbases_per_s": round(batch.n_bases / kern / 1e6, 2),
                          "frac_of_hbm_line": round(32.25 * n_k / (kern * 1e-3) / 6531.6e9, 4), "n_rare": int(rare.numel()),
                          "n_kmers": n_k}), flush=True)
    if args.check:
        ok = all(np.array_equal(v, results["resident"]) for v in results.values()) if "resident" in results else None
        eng.docfreq_mode = "stream"
        ts = eng.count_docfreq(reads, k)
        eng.docfreq_mode = "resident"
        tr = eng.count_docfreq(reads, k)
        ks, rs, ms = (x.cpu().numpy() for x in eng.table_select(ts, 0, 0xFFFFFFFF, 0xFFFFFFFF, with_counts=True))
        kr, rr, mr = (x.cpu().numpy() for x in eng.table_select(tr, 0, 0xFFFFFFFF, 0xFFFFFFFF, with_counts=True))
        os_, or_ = np.argsort(ks.view(np.uint64)), np.argsort(kr.view(np.uint64))
        tab_ok = bool(np.array_equal(ks[os_], kr[or_]) and np.array_equal(rs[os_], rr[or_]) and np.array_equal(ms[os_], mr[or_]))
        print(json.dumps({"rare_equal": ok, "table_equal": tab_ok, "distinct": int(ks.size), "dense": bool(ts.dense)}), flush=True)


if __name__ == "__main__":
    main()
