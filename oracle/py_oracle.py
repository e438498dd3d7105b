"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the unique-k-mer recruitment path.

A plain-Python restatement of the reference algorithm, in the reference's own
string domain, written as order-independent closed forms so that it can be
compared with the device path without depending on dict/set iteration order.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
leg may import it; nothing under ``centroflye_b200/`` does.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(/root/reference/scripts, with a stub ``Bio.SeqIO``) on seeded synthetic NCRF
reports and commits its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against them.

Reference lines followed (relative to /root/reference/scripts):
  segment_units        ncrf_parser.py:28-59
  kmer_freqs           distance_based_kmer_recruitment.py:39-63
  rare_kmers           distance_based_kmer_recruitment.py:66-82
  reads_kmer_clouds    read_kmer_cloud.py:18-40
  filter_clouds        read_kmer_cloud.py:43-54
  dist_counts          distance_based_kmer_recruitment.py:85-128
  filter_edges         distance_based_kmer_recruitment.py:131-149
  result_files         distance_based_kmer_recruitment.py:152-171
"""
import itertools
import math
import re
from collections import Counter, defaultdict


def segment_units(r_al, m_al, motif, n=1):
    """[(start, end)] alignment-column intervals of the units (ncrf_parser.py:34-53)."""
    pattern = re.compile("".join(f"{b}-*" for b in motif) * n)
    matches = list(pattern.finditer(m_al.upper()))
    if not matches:
        return []
    coords = [m.start() for m in matches] + [matches[-1].end()]
    if coords[0] > len(motif) * 0.2:
        coords.insert(0, 0)
    if coords[-1] < len(r_al) - len(motif) * 0.2:
        coords.append(len(r_al))
    return list(zip(coords[:-1], coords[1:]))


def kmer_freqs(records, k, max_nonuniq):
    """kmer -> number of records containing it, for k-mers that occur more than once in at
    most ``max_nonuniq`` records.  Closed form of the sequential update at :55-62: the
    running ``non_unique_freqs`` only grows, so a k-mer survives iff its FINAL value is
    <= max_nonuniq, and while it survives every containing record adds one."""
    n_reads, n_multi = Counter(), Counter()
    for rec in records.values():
        s = rec.r_al.replace("-", "")
        per_read = Counter(s[i:i + k] for i in range(len(s) - k + 1))
        for kmer, c in per_read.items():
            n_reads[kmer] += 1
            if c > 1:
                n_multi[kmer] += 1
    return {kmer: c for kmer, c in n_reads.items() if n_multi[kmer] <= max_nonuniq}


def rare_band(bottom, top, coverage, kmer_survival_rate):
    # same evaluation order as :74-75 (float64)
    return bottom * coverage * kmer_survival_rate, top * coverage * kmer_survival_rate


def rare_kmers(records, k, bottom, top, coverage, kmer_survival_rate, max_nonuniq):
    left, right = rare_band(bottom, top, coverage, kmer_survival_rate)
    return {kmer for kmer, f in kmer_freqs(records, k, max_nonuniq).items() if left <= f <= right}


def reads_kmer_clouds(records, n, k, genomic_kmers):
    """r_id -> list (one per unit) of sets of genomic k-mers lying wholly inside the unit."""
    out = {}
    for r_id, rec in records.items():
        clouds = []
        for st, en in segment_units(rec.r_al, rec.m_al, rec.motif, n=n):
            u = rec.r_al[st:en].upper().replace("-", "")
            clouds.append({u[i:i + k] for i in range(len(u) - k + 1) if u[i:i + k] in genomic_kmers})
        out[r_id] = clouds
    return out


def filter_clouds(clouds, min_mult=2, max_mult=math.inf):
    mult = Counter(kmer for cl in clouds.values() for unit in cl for kmer in unit)
    return {r_id: [{kmer for kmer in unit if max_mult >= mult[kmer] >= min_mult} for unit in cl]
            for r_id, cl in clouds.items()}


def dist_counts(clouds, min_n, max_n, min_d, max_d):
    """(a, b, d) -> count over reads [min_n, max_n) of unit pairs (i, i+d) with a in unit i,
    b in unit i+d, a != b.  ``kmer_clouds[:-dist]`` at :121 makes dist = 0 contribute nothing."""
    cnt = defaultdict(int)
    for _, units in itertools.islice(clouds.items(), min_n, max_n):
        for d in range(max(min_d, 1), max_d + 1):
            for i in range(len(units) - d):
                for a in units[i]:
                    for b in units[i + d]:
                        if a != b:
                            cnt[(a, b, d)] += 1
    return cnt


def filter_edges(cnt, min_coverage, rel_threshold=0.8):
    total = defaultdict(int)
    for (a, b, _), c in cnt.items():
        total[(a, b)] += c
    edges = {(d, a, b, c) for (a, b, d), c in cnt.items()
             if c >= min_coverage and c / total[(a, b)] >= rel_threshold}
    selected = {e[1] for e in edges} | {e[2] for e in edges}
    return selected, edges


def recruit(records, k=19, coverage=32, min_coverage=4, min_n=0, max_n=None, min_d=1, max_d=150,
            bottom=0.9, top=3.0, kmer_survival_rate=0.34, max_nonuniq=3):
    """main() of the reference script, :174-208, returning everything along the way."""
    rare = rare_kmers(records, k, bottom, top, coverage, kmer_survival_rate, max_nonuniq)
    clouds = reads_kmer_clouds(records, 1, k, rare)
    cnt = dist_counts(clouds, min_n, max_n, min_d, max_d)
    selected, edges = filter_edges(cnt, min_coverage)
    return {"rare": rare, "clouds": clouds, "n_increments": sum(cnt.values()), "n_keys": len(cnt),
            "selected": selected, "edges": edges}


def result_files(selected, edges):
    """Contents of the two output files (:158-171) with the hash-seed dependent edge order
    canonicalised by sorting the lines."""
    kmers_txt = "".join(kmer + "\n" for kmer in sorted(selected))
    edge_lines = sorted(f"{d} {a} {b} {c}\n" for d, a, b, c in edges)
    return kmers_txt, edge_lines
