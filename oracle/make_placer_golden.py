#!/usr/bin/env python
"""read_positions.csv of the UNMODIFIED reference read_placer.py on the golden reports (TEST INFRASTRUCTURE ONLY).

    python oracle/make_placer_golden.py

For every golden case (tests/golden/<case>/report.ncrf.gz + the unique k-mers of its parameter point 0, i.e. what
distance_based_kmer_recruitment.py wrote) it runs scripts/read_placer.py:ReadPlacer(params).run() of the reference with
oracle/bio_shim standing in for Biopython and stores the file it wrote as tests/golden/<case>/read_positions_<tag>.csv
plus the parameters.  Placed reads come in the greedy order (deterministic); the trailing "r_id None" lines of unplaced
reads come in set order (PYTHONHASHSEED-dependent) and are compared as a set.
"""
import argparse
import contextlib
import gzip
import io
import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "bio_shim"))
sys.path.insert(0, os.path.join(REF, "scripts"))

from centroflye_b200.encode import ints_to_kmers  # noqa: E402

VARIANTS = {
    "default": dict(n_motif=1, min_cloud_kmer_freq=2, min_kmer_mult=2, min_unit=2, min_inters=10, prefix_threshold=1000),
    "loose": dict(n_motif=1, min_cloud_kmer_freq=1, min_kmer_mult=2, min_unit=1, min_inters=3, prefix_threshold=500),
}


def main():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import read_placer
    golden = os.path.join(ROOT, "tests", "golden")
    for case in sorted(os.listdir(golden)):
        d = os.path.join(golden, case)
        if not os.path.exists(os.path.join(d, "report.ncrf.gz")):
            continue
        p0 = json.load(open(os.path.join(d, "p0.json")))
        k = p0["params"]["k"]
        sel = np.load(os.path.join(d, "p0.npz"))["selected"]
        with tempfile.TemporaryDirectory() as tmp:
            rep = os.path.join(tmp, "report.ncrf")
            with gzip.open(os.path.join(d, "report.ncrf.gz"), "rb") as f, open(rep, "wb") as g:
                g.write(f.read())
            kfn = os.path.join(tmp, "kmers.txt")
            with open(kfn, "w") as f:
                f.write("".join(kmer + "\n" for kmer in ints_to_kmers(sel, k)))
            for tag, v in VARIANTS.items():
                out = os.path.join(tmp, "out_" + tag)
                params = argparse.Namespace(ncrf=rep, genomic_kmers=kfn, k_cloud=k, outdir=out, **v)
                with contextlib.redirect_stdout(io.StringIO()):
                    read_placer.ReadPlacer(params).run()
                text = open(os.path.join(out, "read_positions.csv")).read()
                with open(os.path.join(d, f"read_positions_{tag}.csv"), "w") as f:
                    f.write(text)
                with open(os.path.join(d, f"read_positions_{tag}.json"), "w") as f:
                    json.dump(dict(v, k_cloud=k, genomic_kmers="p0.npz:selected"), f, indent=1, sort_keys=True)
                lines = text.splitlines()
                placed = [ln for ln in lines if not ln.endswith("None")]
                print(f"{case} {tag}: {len(sel)} k-mers, {len(lines)} lines, {len(placed)} placed", flush=True)


if __name__ == "__main__":
    main()
