#!/usr/bin/env python
"""T0: the unmodified reference (baseline/_ref, made by baseline/make_ref.py) timed stage by stage on one host core.

Called by bench.py (`cpu_baseline_t0`); BENCH INFRASTRUCTURE -- nothing under centroflye_b200/ imports this.
The reference is single-threaded Python: stage A ~0.4-0.8 Mbases/s, stage B ~1.3-2 Mbases/s plus the
2055-group regex of get_motif_alignments, stage C ~2 M increments/s at 40 B per counter -- configs[1] would take hours
and ~1 TB, so every stage runs on a BOUNDED sample of the bench's own reads and reports a rate:

  A  get_kmer_freqs_from_ncrf_report (dbkr.py:39-63) on the first n_a records of the report          -> bases/s
  B  get_reads_kmer_clouds (read_kmer_cloud.py:34-40, incl. NCRF_Record.get_motif_alignments) on the first n_b
     records against the rare set get_rare_kmers-style band of the sample itself                        -> bases/s
  C  get_kmer_dist_map + filter_dist_tuples (dbkr.py:85-149) on those clouds with max_d cut so that the number of
     increments stays bounded                                                                            -> increments/s
"""
import contextlib
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "scripts", "distance_based_kmer_recruitment.py"))


def _import_reference():
    for p in (os.path.join(REF, "bio_shim"), os.path.join(REF, "scripts")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # ncrf_parser.py:74-75: invalid escape sequence '\s'
        dbkr = importlib.import_module("distance_based_kmer_recruitment")
        rkc = importlib.import_module("read_kmer_cloud")
        parser = importlib.import_module("ncrf_parser")
    for m in (dbkr, rkc, parser):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF)), m.__file__
    return dbkr, rkc, parser


def run(report_fn, k, max_nonuniq, min_d, max_d, min_coverage, n_a=60, n_b=12, c_budget=2.0e7):
    """report_fn: an NCRF report holding at least n_a records of the bench workload."""
    dbkr, rkc, parser = _import_reference()
    out = {"kind": "reference", "cores": 1, "source": "baseline/_ref (unmodified scripts/*.py of the reference)"}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        t = time.perf_counter()
        report = parser.NCRF_Report(report_fn)
        out["parse_s"] = time.perf_counter() - t
        recs = list(report.records.items())
        # --- stage A
        rep_a = parser.NCRF_Report.__new__(parser.NCRF_Report)
        rep_a.records = dict(recs[:n_a])
        bases_a = sum(len(r.r_al.replace('-', '')) for r in rep_a.records.values())
        t = time.perf_counter()
        freqs = dbkr.get_kmer_freqs_from_ncrf_report(rep_a, k, False, max_nonuniq)
        dt = time.perf_counter() - t
        out["stage_a"] = {"reads": len(rep_a.records), "bases": bases_a, "seconds": dt, "bases_per_s": bases_a / dt,
                          "distinct_kmers": len(freqs)}
        # --- stage B (the band of the sample itself: k-mers seen in 2..32 of the n_a reads)
        rare = {kmer for kmer, f in freqs.items() if 2 <= f <= 32}
        rep_b = parser.NCRF_Report.__new__(parser.NCRF_Report)
        rep_b.records = dict(recs[:n_b])
        bases_b = sum(len(r.r_al.replace('-', '')) for r in rep_b.records.values())
        t = time.perf_counter()
        clouds = rkc.get_reads_kmer_clouds(rep_b, n=1, k=k, genomic_kmers=rare)
        dt = time.perf_counter() - t
        n_units = sum(len(c.kmers) for c in clouds.values())
        out["stage_b"] = {"reads": len(rep_b.records), "bases": bases_b, "seconds": dt, "bases_per_s": bases_b / dt,
                          "units": n_units, "rare_kmers": len(rare)}
        # --- stage C/D: cut max_d so that the increments stay within budget (closed form of dbkr.py:121-126)
        def increments(n_c, md):
            sizes = [[len(c) for c in kc.kmers] for kc in list(clouds.values())[:n_c]]
            return sum(sz[i] * sz[i + d] for sz in sizes for d in range(max(min_d, 1), md + 1) for i in range(len(sz) - d))
        n_c, md = min(4, len(clouds)), max_d
        while increments(n_c, md) > c_budget and (md > 1 or n_c > 1):
            if md > 1:
                md = max(1, md // 2)
            else:
                n_c -= 1
        n_inc = increments(n_c, md)
        t = time.perf_counter()
        dist_cnt, kmer_index = dbkr.get_kmer_dist_map(clouds, rare, 0, n_c, min_d, md, False)
        sel, edges = dbkr.filter_dist_tuples(dist_cnt, min_coverage)
        dt = time.perf_counter() - t
        out["stage_c"] = {"reads": n_c, "max_d": md, "increments_upper": n_inc, "seconds": dt,
                          "increments_per_s": n_inc / dt if dt > 0 else None, "edges": len(edges)}
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run(sys.argv[1], 19, 3, 1, 150, 4), indent=1))
