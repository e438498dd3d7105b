#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Runs only in the build container, where /root/reference exists:

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py dxz1_small # one case

For every case it
  1. simulates a tandem-repeat genome with the reference's own
     scripts/simulate_tandem_repeat.py:generate_mutations (legacy np.random, seeded),
  2. simulates reads + the truth NCRF report with centroflye_b200.synth,
  3. imports the reference modules (scripts/ncrf_parser.py, read_kmer_cloud.py,
     distance_based_kmer_recruitment.py) with oracle/bio_shim on the path and calls
     the reference functions exactly as its main() does (dbkr.py:174-208),
  4. stores the report and every intermediate (P1..P6 of SURVEY.md §8c) under
     tests/golden/<case>/.

The fixtures travel to the GPU box; the reference does not.
"""
import gzip
import hashlib
import io
import json
import os
import sys
import tempfile
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "bio_shim"))
sys.path.insert(0, os.path.join(REF, "scripts"))

from centroflye_b200 import synth  # noqa: E402
from centroflye_b200.encode import kmers_to_ints  # noqa: E402

CASES = {
    # name: genome + read model + recruitment parameter points
    "dxz1_small": dict(
        unit_fasta="supplementary_data/DXZ1_rc.fasta", multiplicity=12, div_rate=0.01, genome_seed=1,
        flank_len=3000, coverage=14, error_rate=0.06, read_seed=2, median_len=9000, sigma=0.4,
        min_len=5200, max_len=30000,
        points=[dict(k=19, coverage=14), dict(k=19, coverage=14, max_nonuniq=0, min_coverage=2),
                dict(k=15, coverage=14, bottom=0.5, top=2.0), dict(k=27, coverage=10, min_coverage=3),
                dict(k=19, coverage=14, min_coverage=8), dict(k=19, coverage=14, max_nonuniq=0)],
    ),
    "rand311": dict(
        unit_random=(311, 7), multiplicity=50, div_rate=0.02, genome_seed=3,
        flank_len=2000, coverage=11, error_rate=0.05, read_seed=4, median_len=8000, sigma=0.5,
        min_len=5200, max_len=40000,
        points=[dict(k=19, coverage=11), dict(k=11, coverage=11, max_distance=7),
                dict(k=23, coverage=11, min_nreads=3, max_nreads=15, min_coverage=2),
                dict(k=31, coverage=9, min_distance=2, max_distance=20),
                dict(k=19, coverage=11, min_distance=0, max_distance=3, kmer_survival_rate=0.4)],
    ),
    "d6z1_noisy": dict(
        unit_fasta="supplementary_data/D6Z1.fasta", multiplicity=8, div_rate=0.01, genome_seed=4,
        flank_len=2500, coverage=20, error_rate=0.12, read_seed=5, median_len=9000, sigma=0.4,
        min_len=5200, max_len=30000,
        points=[dict(k=19, coverage=20, kmer_survival_rate=0.09), dict(k=13, coverage=20, kmer_survival_rate=0.2)],
    ),
}

DEFAULTS = dict(min_coverage=4, min_nreads=0, max_nreads=sys.maxsize, min_distance=1, max_distance=150,
                bottom=0.9, top=3.0, kmer_survival_rate=0.34, max_nonuniq=3)


def _ref():
    with contextlib.redirect_stderr(io.StringIO()):  # SyntaxWarning noise from ncrf_parser.py:74-75
        import warnings
        warnings.simplefilter("ignore")
        import distance_based_kmer_recruitment as dbkr
        import ncrf_parser
        import read_kmer_cloud
        import simulate_tandem_repeat
        from utils.bio import read_bio_seq
    return dbkr, ncrf_parser, read_kmer_cloud, simulate_tandem_repeat, read_bio_seq


NEW_ONLY = False  # --new-only: keep the fixtures that exist, add the missing ones


def _ref_get_kmer_counts_reads():
    """get_kmer_counts_reads compiled from its own source lines in the reference file (nothing else of the module)."""
    import ast
    path = os.path.join(REF, "scripts", "better_consensus_unit_reconstruction.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_kmer_counts_reads")
    ns = {}
    exec(compile(ast.Module(body=[ast.parse("from collections import defaultdict").body[0], fn], type_ignores=[]),
                 path, "exec"), ns)
    return ns["get_kmer_counts_reads"]


def _dump(path, obj):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, sort_keys=True, separators=(",", ":")).encode())


def _clouds_to_csr(clouds, order, k, tag):
    """dict r_id -> list of sorted k-mer lists  ->  units-per-read, unit_ptr, u64 k-mers."""
    n_units = np.array([len(clouds[r]) for r in order], dtype=np.int64)
    sizes = np.array([len(u) for r in order for u in clouds[r]], dtype=np.int64)
    flat = [kmer for r in order for u in clouds[r] for kmer in u]
    ptr = np.zeros(sizes.size + 1, dtype=np.int64)
    np.cumsum(sizes, out=ptr[1:])
    vals = kmers_to_ints(flat, k)
    for lo, hi in zip(ptr[:-1], ptr[1:]):
        vals[lo:hi].sort()
    return {f"{tag}_units_per_read": n_units, f"{tag}_unit_ptr": ptr, f"{tag}_kmers": vals}


def build_case(name, spec, outdir):
    dbkr, ncrf_parser, rkc, sim, read_bio_seq = _ref()
    os.makedirs(outdir, exist_ok=True)
    if "unit_fasta" in spec:
        unit = read_bio_seq(os.path.join(REF, spec["unit_fasta"])).upper()
    else:
        unit = synth.random_unit(*spec["unit_random"])
    np.random.seed(spec["genome_seed"])
    tr, _, flanked, _ = sim.generate_mutations(unit, spec["multiplicity"], spec["div_rate"],
                                               flank_len=spec["flank_len"])
    from centroflye_b200.encode import ascii_to_codes
    genome = ascii_to_codes(flanked)
    reads = synth.simulate_reads(genome, spec["flank_len"], len(tr), unit, spec["coverage"],
                                 spec["error_rate"], spec["read_seed"], median_len=spec["median_len"],
                                 sigma=spec["sigma"], min_len=spec["min_len"], max_len=spec["max_len"])
    report_path = os.path.join(outdir, "report.ncrf")
    # exercise ncrf_parser.py:91-93: a shorter duplicate alignment of read 0 before and after the
    # real one, a too-short record, comments and blank lines
    with open(report_path, "w") as f:
        f.write("# synthetic NCRF report\n\n")
        for i, rd in enumerate(reads):
            l1, l2 = synth.ncrf_record_lines(rd, unit)
            if i in (0, 3):
                cut = rd.r_row.size * 2 // 3
                short = synth.SynthRead(rd.r_id, rd.r_len, rd.r_st, rd.r_st + int((rd.r_row[:cut] != 4).sum()),
                                        "+", rd.r_row[:cut], rd.m_row[:cut], rd.unit_cols[:0])
                s1, s2 = synth.ncrf_record_lines(short, unit)
                if i == 0:
                    f.write(s1 + "\n" + s2 + "\n\n")
                    f.write(l1 + "\n" + l2 + "\n\n")
                else:
                    f.write(l1 + "\n" + l2 + "\n# duplicate follows\n" + s1 + "\n" + s2 + "\n\n")
            else:
                f.write(l1 + "\n" + l2 + "\n\n")
        tiny = synth.SynthRead("tiny_read", 900, 0, 800, "+", reads[0].r_row[:820], reads[0].m_row[:820],
                               reads[0].unit_cols[:0])
        t1, t2 = synth.ncrf_record_lines(tiny, unit)
        f.write(t1 + "\n" + t2 + "\n")

    report = ncrf_parser.NCRF_Report(report_path)
    meta = {"case": name, "motif_len": len(unit), "n_records": len(report.records),
            "n_bases": sum(len(r.r_al.replace("-", "")) for r in report.records.values()),
            "reference_commit": "b2a4378bc254cc59bf13ae4e802dab478666e079"}
    parsed = {r_id: dict(r_len=r.r_len, r_al_len=r.r_al_len, r_st=r.r_st, r_en=r.r_en, strand=r.strand,
                         m_al_len=r.m_al_len, al_score=r.al_score,
                         r_al_md5=hashlib.md5(r.r_al.encode()).hexdigest(),
                         m_al_md5=hashlib.md5(r.m_al.encode()).hexdigest())
              for r_id, r in report.records.items()}
    seg = {}
    for n in (1, 2):
        seg[str(n)] = {r_id: [[ma.start, ma.end] for ma in r.get_motif_alignments(n=n)]
                       for r_id, r in report.records.items()}
    _dump(os.path.join(outdir, "parsed.json.gz"),
          {"meta": meta, "order": list(report.records.keys()), "records": parsed, "segments": seg,
           "discarded": sorted(report.discarded_reads),
           "classify_3000": [sorted(x) for x in report.classify(large_threshold=3000)]})

    for k_count in (19, 30):  # total-occurrence counts of the reference's get_kmer_counts_reads (row f3)
        fn = os.path.join(outdir, f"kmer_counts_k{k_count}.json")
        if not (NEW_ONLY and os.path.exists(fn)):
            counts = _ref_get_kmer_counts_reads()(report, k=k_count)
            lines = "".join(f"{kmer} {c}\n" for kmer, c in sorted(counts.items()))
            with open(fn, "w") as f:
                json.dump({"k": k_count, "n_distinct": len(counts), "n_total": int(sum(counts.values())),
                           "max_count": int(max(counts.values())), "sorted_lines_md5": hashlib.md5(lines.encode()).hexdigest(),
                           "source": "scripts/better_consensus_unit_reconstruction.py:127-135, the function's own source "
                                     "compiled out of the reference file (the module imports edlib, absent here)"},
                          f, indent=1, sort_keys=True)
    for pi, point in enumerate(spec["points"]):
        if NEW_ONLY and os.path.exists(os.path.join(outdir, f"p{pi}.json")):
            continue
        p = dict(DEFAULTS)
        p.update(point)
        k = p["k"]
        all_kmers = dbkr.get_kmer_freqs_from_ncrf_report(report, k=k, verbose=False, max_nonuniq=p["max_nonuniq"])
        rare = dbkr.get_rare_kmers(report, k=k, bottom=p["bottom"], top=p["top"], coverage=p["coverage"],
                                   kmer_survival_rate=p["kmer_survival_rate"], max_nonuniq=p["max_nonuniq"],
                                   verbose=False)
        clouds = rkc.get_reads_kmer_clouds(report, n=1, k=k, genomic_kmers=rare)
        p3 = {r_id: [sorted(u) for u in c.kmers] for r_id, c in clouds.items()}
        dist_cnt, kmer_index = dbkr.get_kmer_dist_map(clouds, rare, min_n=p["min_nreads"], max_n=p["max_nreads"],
                                                      min_d=p["min_distance"], max_d=p["max_distance"],
                                                      verbose=False)
        n_incr = sum(sum(dd.values()) for dt in dist_cnt.values() for dd in dt)
        n_keys = sum(len(dd) for dt in dist_cnt.values() for dd in dt)
        uniq, edges = dbkr.filter_dist_tuples(dist_cnt, min_coverage=p["min_coverage"])
        with tempfile.TemporaryDirectory() as td:
            dbkr.output_results(kmer_index=kmer_index, min_coverage=p["min_coverage"], unique_kmers_ind=uniq,
                                dist_edges=edges, outdir=td)
            kmers_txt = open(os.path.join(td, f"unique_kmers_min_edge_cov_{p['min_coverage']}.txt")).read()
            edge_lines = sorted(open(os.path.join(td, f"unique_edges_min_edge_cov_{p['min_coverage']}.txt")).readlines())
        # P6 on clouds re-built the way read_placer.py:106-114 does (plain set[str] of recruited k-mers)
        recruited = set(kmers_txt.split())
        clouds2 = rkc.get_reads_kmer_clouds(report, n=1, k=k, genomic_kmers=recruited)
        clouds2 = rkc.filter_reads_kmer_clouds(clouds2, min_mult=2)
        p6 = {r_id: [sorted(u) for u in c.kmers] for r_id, c in clouds2.items()}
        clouds_n2 = rkc.get_reads_kmer_clouds(report, n=2, k=k, genomic_kmers=rare)
        p3n2 = {r_id: [sorted(u) for u in c.kmers] for r_id, c in clouds_n2.items()}

        keys = kmers_to_ints(list(all_kmers.keys()), k)
        vals = np.fromiter(all_kmers.values(), dtype=np.uint32, count=len(all_kmers))
        order = np.argsort(keys)
        arrays = dict(all_keys=keys[order], all_counts=vals[order],
                      rare=np.sort(kmers_to_ints(sorted(rare), k)),
                      selected=np.sort(kmers_to_ints(kmers_txt.split(), k)))
        for tag, cl in (("clouds", p3), ("clouds_n2", p3n2), ("placer", p6)):
            arrays.update(_clouds_to_csr(cl, list(report.records.keys()), k, tag))
        e = [ln.split() for ln in edge_lines]
        e_d = np.array([int(x[0]) for x in e], dtype=np.int32)
        e_a = kmers_to_ints([x[1] for x in e], k)
        e_b = kmers_to_ints([x[2] for x in e], k)
        e_c = np.array([int(x[3]) for x in e], dtype=np.uint32)
        eo = np.lexsort((e_b, e_a, e_d))
        arrays.update(edge_d=e_d[eo], edge_a=e_a[eo], edge_b=e_b[eo], edge_cnt=e_c[eo])
        np.savez_compressed(os.path.join(outdir, f"p{pi}.npz"), **arrays)
        if p["max_nreads"] == sys.maxsize:
            p = dict(p, max_nreads=None)
        with open(os.path.join(outdir, f"p{pi}.json"), "w") as f:
            json.dump({"params": p, "n_all_kmers": len(all_kmers), "n_rare": len(rare),
                       "n_increments": int(n_incr), "n_keys": int(n_keys), "n_edges": len(edge_lines),
                       "n_selected": len(kmers_txt.split()),
                       "unique_kmers_txt_md5": hashlib.md5(kmers_txt.encode()).hexdigest(),
                       "edge_lines_sorted_md5": hashlib.md5("".join(edge_lines).encode()).hexdigest()},
                      f, indent=1, sort_keys=True)
        print(f"{name} point {pi}: k={k} all={len(all_kmers)} rare={len(rare)} incr={n_incr} keys={n_keys} "
              f"edges={len(edge_lines)} unique={len(kmers_txt.split())}", flush=True)
    with open(report_path, "rb") as f, gzip.GzipFile(report_path + ".gz", "wb", mtime=0) as g:
        g.write(f.read())
    os.remove(report_path)
    print(f"{name}: {meta}", flush=True)


def main():
    global NEW_ONLY
    NEW_ONLY = "--new-only" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(CASES)
    for name in names:
        build_case(name, CASES[name], os.path.join(ROOT, "tests", "golden", name))


if __name__ == "__main__":
    main()
