#!/usr/bin/env python
"""Read recruitment pre-filter: the device kernel against the reference binary (oracle/_ref/rr) on synthetic reads.
    python tools/rr_bench.py [--mbases 300] [--ref-mbases 20]
90 % of the reads are random sequence (whole-genome reads that do not hit the centromere: both strands are scanned to
the end), 10 % carry DXZ1 copies with 10-15 % errors (kept, the scan stops early).  Threshold 350 (run_read_recruitment.sh)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbases", type=float, default=300.0)
    ap.add_argument("--ref-mbases", type=float, default=20.0)
    ap.add_argument("--threshold", type=int, default=350)
    args = ap.parse_args()
    import torch
    from centroflye_b200 import read_recruitment as rr, synth
    from centroflye_b200.engine import default_engine
    rng = np.random.default_rng(0)
    unit = synth.load_genome("cenx_dxz1_m1500_s1")[3]
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, total = [], 0
    while total < args.mbases * 1e6:
        n = int(np.clip(np.exp(rng.normal(np.log(20000), 0.5)), 2000, 200000))
        codes = rng.integers(0, 4, size=n, dtype=np.uint8)
        if rng.random() < 0.1:  # a few noisy unit copies somewhere inside
            u = np.frombuffer(unit.encode(), dtype=np.uint8).copy()
            hit = rng.random(u.size) < 0.12
            u[hit] = lut[rng.integers(0, 4, size=int(hit.sum()))]
            at = int(rng.integers(0, max(1, n - u.size)))
            s = lut[codes].copy()
            s[at:at + u.size] = u[: max(0, min(u.size, n - at))]
            seqs.append(s.tobytes().decode())
        else:
            seqs.append(lut[codes].tobytes().decode())
        total += n
    eng = default_engine()
    rr.recruit(unit, seqs[:50], args.threshold)  # warm-up
    torch.cuda.synchronize()
    t = time.perf_counter()
    eng.events = []
    keep = rr.recruit(unit, seqs, args.threshold)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t
    kernel_ms = eng.stage_times_ms().get("rr_filter", 0.0)
    eng.events = None
    # the kernel alone: the launch is inside recruit(); time it again through the stage events of the engine
    out = {"reads": len(seqs), "bases": total, "kept": int(keep.sum()), "threshold": args.threshold,
           "gpu_wall_s_incl_host_packing_and_h2d": wall, "gpu_bases_per_s_wall": total / wall,
           "kernel_ms": kernel_ms, "kernel_bases_per_s": total / (kernel_ms * 1e-3) if kernel_ms else None}
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rr")):
        n_ref, b_ref = 0, 0
        while n_ref < len(seqs) and b_ref < args.ref_mbases * 1e6:
            b_ref += len(seqs[n_ref])
            n_ref += 1
        with tempfile.TemporaryDirectory() as tmp:
            with open(tmp + "/unit.fasta", "w") as f:
                f.write(">u\n" + unit + "\n")
            with open(tmp + "/reads.fasta", "w") as f:
                f.write("".join(f">r{i}\n{s}\n" for i, s in enumerate(seqs[:n_ref])))
            t = time.perf_counter()
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "rr"), tmp + "/unit.fasta", tmp + "/reads.fasta",
                                   tmp + "/out.fasta", str(args.threshold)])
            dt = time.perf_counter() - t
            kept_ref = open(tmp + "/out.fasta").read().count(">")
        out.update({"reference_binary": {"reads": n_ref, "bases": b_ref, "seconds": dt, "bases_per_s_1_core": b_ref / dt,
                                         "kept": kept_ref, "kept_equal": kept_ref == int(keep[:n_ref].sum())}})
        out["speedup_vs_1_core"] = out["gpu_bases_per_s_wall"] / (b_ref / dt)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
