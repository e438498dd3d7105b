"""GPU parity tests (run with ``-m gpu`` on the B200 box).  Every check goes through the drop-in
modules, i.e. through the C ABI of libcfk.so, and compares bit-for-bit with what the UNMODIFIED
reference produced for the same report (tests/golden/, made by oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import all_case_points, clouds_to_csr, golden_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import centroflye_b200.distance_based_kmer_recruitment as dbkr
    import centroflye_b200.read_kmer_cloud as rkc
    from centroflye_b200.ncrf_parser import NCRF_Report
    return dbkr, rkc, NCRF_Report


def _csr_of(clouds, k):
    """CloudDict -> (units_per_read, unit_ptr, u64 k-mers sorted per unit) straight from the device arrays."""
    st = clouds.device_state()
    assert st is not None
    unit_ptr, ids, keys = st.host()
    return np.diff(st.read_unit_ptr), unit_ptr, keys[ids]


@pytest.fixture(params=["auto", "exact"])
def pair_mode(request):
    """Stage C flavour: auto = sketch kernel where min_coverage allows it, exact = always the exact tables."""
    from centroflye_b200.engine import default_engine
    eng = default_engine()
    old, eng.pair_mode = eng.pair_mode, request.param
    yield request.param
    eng.pair_mode = old


@pytest.mark.parametrize("case,pi", all_case_points())
def test_recruitment_matches_reference(golden, mods, case, pi, tmp_path, pair_mode):
    dbkr, rkc, NCRF_Report = mods
    g = golden(case)
    meta, arr = g.point(pi)
    p = meta["params"]
    k = p["k"]
    max_n = p["max_nreads"] if p["max_nreads"] is not None else 2**63 - 1
    rep = NCRF_Report(g.report_path)
    order = list(rep.records.keys())

    # P1: document frequencies (dbkr.py:39-63)
    freqs = dbkr.get_kmer_freqs_from_ncrf_report(rep, k=k, verbose=False, max_nonuniq=p["max_nonuniq"])
    assert np.array_equal(freqs.keys_u64, arr["all_keys"])
    assert np.array_equal(freqs.counts, arr["all_counts"])

    # P2: rare band (dbkr.py:66-82)
    rare = dbkr.get_rare_kmers(rep, k=k, bottom=p["bottom"], top=p["top"], coverage=p["coverage"],
                               kmer_survival_rate=p["kmer_survival_rate"], max_nonuniq=p["max_nonuniq"],
                               verbose=False)
    from centroflye_b200.encode import kmers_to_ints
    assert np.array_equal(np.sort(kmers_to_ints(sorted(rare), k)), arr["rare"])

    # P3: clouds (read_kmer_cloud.py:18-40), n = 1 and n = 2
    clouds = rkc.get_reads_kmer_clouds(rep, n=1, k=k, genomic_kmers=rare)
    assert list(clouds.keys()) == order
    upr, ptr, vals = _csr_of(clouds, k)
    assert np.array_equal(upr, arr["clouds_units_per_read"])
    assert np.array_equal(ptr, arr["clouds_unit_ptr"])
    assert np.array_equal(vals, arr["clouds_kmers"])
    upr2, ptr2, vals2 = _csr_of(rkc.get_reads_kmer_clouds(rep, n=2, k=k, genomic_kmers=rare), k)
    assert np.array_equal(upr2, arr["clouds_n2_units_per_read"])
    assert np.array_equal(ptr2, arr["clouds_n2_unit_ptr"])
    assert np.array_equal(vals2, arr["clouds_n2_kmers"])

    # P4/P5: distance graph + edge filter (dbkr.py:85-149) and the two files (dbkr.py:152-171)
    dist_cnt, kmer_index = dbkr.get_kmer_dist_map(clouds, rare, min_n=p["min_nreads"], max_n=max_n,
                                                  min_d=p["min_distance"], max_d=p["max_distance"], verbose=False)
    uniq, edges = dbkr.filter_dist_tuples(dist_cnt, min_coverage=p["min_coverage"])
    assert dist_cnt.n_increments == meta["n_increments"]
    assert len(edges) == meta["n_edges"]
    keys = kmer_index.keys_u64
    assert np.array_equal(edges.dist, arr["edge_d"])
    assert np.array_equal(keys[edges.i], arr["edge_a"])
    assert np.array_equal(keys[edges.j], arr["edge_b"])
    assert np.array_equal(edges.freq, arr["edge_cnt"])
    assert np.array_equal(keys[np.array(sorted(uniq), dtype=np.int64)], arr["selected"])
    dbkr.output_results(kmer_index=kmer_index, min_coverage=p["min_coverage"], unique_kmers_ind=uniq,
                        dist_edges=edges, outdir=str(tmp_path))
    kmers_txt = open(tmp_path / f"unique_kmers_min_edge_cov_{p['min_coverage']}.txt").read()
    assert hashlib.md5(kmers_txt.encode()).hexdigest() == meta["unique_kmers_txt_md5"]
    edge_lines = sorted(open(tmp_path / f"unique_edges_min_edge_cov_{p['min_coverage']}.txt").readlines())
    assert hashlib.md5("".join(edge_lines).encode()).hexdigest() == meta["edge_lines_sorted_md5"]

    # P6: what read_placer.py:20-25,106-114 does: plain set[str] of recruited k-mers, then the multiplicity filter
    recruited = set(kmers_txt.split())
    placer = rkc.filter_reads_kmer_clouds(rkc.get_reads_kmer_clouds(rep, n=1, k=k, genomic_kmers=recruited), min_mult=2)
    _, ptr6, vals6 = _csr_of(placer, k)
    assert np.array_equal(ptr6, arr["placer_unit_ptr"])
    assert np.array_equal(vals6, arr["placer_kmers"])


def test_cli_end_to_end(golden, mods, tmp_path):
    """The command centroFlye.py:172-188 runs, byte-for-byte output check."""
    dbkr, _, _ = mods
    g = golden("dxz1_small")
    meta, _ = g.point(0)
    p = meta["params"]
    out = tmp_path / "recruited_unique_kmers"
    dbkr.main(["--ncrf", g.report_path, "--coverage", str(p["coverage"]), "--min-coverage", str(p["min_coverage"]),
               "--outdir", str(out)])
    mc = p["min_coverage"]
    kmers_txt = open(out / f"unique_kmers_min_edge_cov_{mc}.txt").read()
    assert hashlib.md5(kmers_txt.encode()).hexdigest() == meta["unique_kmers_txt_md5"]
    lines = sorted(open(out / f"unique_edges_min_edge_cov_{mc}.txt").readlines())
    assert hashlib.md5("".join(lines).encode()).hexdigest() == meta["edge_lines_sorted_md5"]


def test_materialised_views_and_host_sets(golden, mods):
    """The Python-object side of the boundary: .kmers as list[set[str]], dist_cnt[d][i][j], host-held sets."""
    dbkr, rkc, NCRF_Report = mods
    from oracle import py_oracle
    g = golden("rand311")
    meta, arr = g.point(4)  # min_distance = 0, max_distance = 3: small graph
    p = meta["params"]
    k = p["k"]
    rep = NCRF_Report(g.report_path)
    rare = dbkr.get_rare_kmers(rep, k=k, bottom=p["bottom"], top=p["top"], coverage=p["coverage"],
                               kmer_survival_rate=p["kmer_survival_rate"], max_nonuniq=p["max_nonuniq"], verbose=False)
    want_rare = py_oracle.rare_kmers(rep.records, k, p["bottom"], p["top"], p["coverage"], p["kmer_survival_rate"],
                                     p["max_nonuniq"])
    assert set(rare) == want_rare
    clouds = rkc.get_reads_kmer_clouds(rep, n=1, k=k, genomic_kmers=rare)
    want_clouds = py_oracle.reads_kmer_clouds(rep.records, 1, k, want_rare)
    dist_cnt, kmer_index = dbkr.get_kmer_dist_map(clouds, rare, 0, 2**63 - 1, p["min_distance"], p["max_distance"], False)
    # full counter table, read the way the reference's filter reads it
    want_cnt = py_oracle.dist_counts(want_clouds, 0, None, p["min_distance"], p["max_distance"])
    got = {}
    for d, tables in dist_cnt.items():
        for i, row in enumerate(tables):
            for j, c in row.items():
                got[(i, j, d)] = c
    rev = kmer_index.kmer_of(np.arange(len(kmer_index)))
    assert {(rev[a], rev[b], d): c for (a, b, d), c in got.items()} == dict(want_cnt)
    assert dist_cnt.n_increments == sum(want_cnt.values())
    # materialised sets equal the oracle's, and all_kmers is their concatenation
    for r_id, units in want_clouds.items():
        assert clouds[r_id].kmers == units
        assert sorted(clouds[r_id].all_kmers) == sorted(km for u in units for km in u)
    # after looking at .kmers the host copy is authoritative: mutate it and recount from the sets
    first = next(iter(clouds))
    clouds[first].kmers[0] = set()
    want_clouds[first][0] = set()
    dist2, index2 = dbkr.get_kmer_dist_map(clouds, rare, 0, 2**63 - 1, 1, 3, False)
    uniq2, edges2 = dbkr.filter_dist_tuples(dist2, min_coverage=p["min_coverage"])
    sel, want_edges = py_oracle.filter_edges(py_oracle.dist_counts(want_clouds, 0, None, 1, 3), p["min_coverage"])
    rev2 = index2.kmer_of(np.arange(len(index2)))
    assert {(d, rev2[i], rev2[j], c) for d, i, j, c in edges2} == want_edges
    assert {rev2[i] for i in uniq2} == sel
    # filter on host-held sets writes back in place (read_kmer_cloud.py:49-53)
    ret = rkc.filter_reads_kmer_clouds(clouds, min_mult=3, max_mult=9)
    assert ret is clouds
    want_f = py_oracle.filter_clouds(want_clouds, 3, 9)
    for r_id, units in want_f.items():
        assert clouds[r_id].kmers == units


def test_stage_edge_cases(mods, tmp_path):
    """Empty and degenerate inputs: reads shorter than k, no rare k-mers, no units, max_d < min_d."""
    dbkr, rkc, NCRF_Report = mods
    from centroflye_b200 import synth
    unit = synth.random_unit(97, 11)
    genome, a0, alen = synth.simulate_genome(unit, 400, 0.02, 5, flank_len=500)
    reads = synth.simulate_reads(genome, a0, alen, unit, 6, 0.03, 6, median_len=6000, sigma=0.2, min_len=5300,
                                 max_len=9000)
    path = tmp_path / "r.ncrf"
    synth.write_ncrf_report(path, reads, unit)
    rep = NCRF_Report(str(path))
    assert len(rep.records) > 3
    # band that selects nothing
    rare = dbkr.get_rare_kmers(rep, k=19, bottom=50.0, top=60.0, coverage=30, kmer_survival_rate=1.0, max_nonuniq=3,
                               verbose=False)
    assert len(rare) == 0
    clouds = rkc.get_reads_kmer_clouds(rep, n=1, k=19, genomic_kmers=rare)
    assert all(all(len(u) == 0 for u in c.kmers) for c in clouds.values())
    clouds = rkc.get_reads_kmer_clouds(rep, n=1, k=19, genomic_kmers=rare)
    dist_cnt, idx = dbkr.get_kmer_dist_map(clouds, rare, 0, 2**63 - 1, 1, 150, False)
    uniq, edges = dbkr.filter_dist_tuples(dist_cnt, 4)
    assert uniq == set() and len(edges) == 0
    # inverted distance range, and a read window past the end
    rare = dbkr.get_rare_kmers(rep, k=19, bottom=0.9, top=3.0, coverage=6, kmer_survival_rate=0.5, max_nonuniq=3,
                               verbose=False)
    clouds = rkc.get_reads_kmer_clouds(rep, n=1, k=19, genomic_kmers=rare)
    for args in ((0, 2**63 - 1, 5, 4), (10**6, 2**63 - 1, 1, 150), (2, 2, 1, 150)):
        dist_cnt, _ = dbkr.get_kmer_dist_map(clouds, rare, *args, False)
        uniq, edges = dbkr.filter_dist_tuples(dist_cnt, 1)
        assert uniq == set() and len(edges) == 0
    # n larger than any read has copies: no units at all
    empty = rkc.get_reads_kmer_clouds(rep, n=500, k=19, genomic_kmers=rare)
    assert all(len(c.kmers) == 0 for c in empty.values())
    with pytest.raises(TypeError):
        rkc.get_reads_kmer_clouds(rep, n=1, k=19)
    with pytest.raises(ValueError):
        dbkr.get_rare_kmers(rep, k=32, bottom=0.9, top=3.0, coverage=6, kmer_survival_rate=0.5, max_nonuniq=3,
                            verbose=False)


def test_non_acgt_is_rejected_loudly(mods, tmp_path):
    dbkr, _, NCRF_Report = mods
    from centroflye_b200 import synth
    unit = synth.random_unit(97, 11)
    genome, a0, alen = synth.simulate_genome(unit, 400, 0.02, 5, flank_len=500)
    reads = synth.simulate_reads(genome, a0, alen, unit, 3, 0.03, 6, median_len=6000, sigma=0.2, min_len=5300,
                                 max_len=9000)
    path = tmp_path / "r.ncrf"
    synth.write_ncrf_report(path, reads, unit)
    text = open(path).read().split("\n")
    i = next(i for i, ln in enumerate(text) if ln.startswith("read_") and int(ln.split()[2][:-2]) >= 5000)
    head, row = text[i].rsplit(" ", 1)
    text[i] = head + " " + row[:100] + "N" + row[101:]
    open(path, "w").write("\n".join(text))
    with pytest.raises(ValueError):
        dbkr.get_rare_kmers(NCRF_Report(str(path)), k=19, bottom=0.9, top=3.0, coverage=6, kmer_survival_rate=0.5,
                            max_nonuniq=3, verbose=False)


def test_sharded_recruitment_two_gpus():
    """One process per GPU over NCCL: sharded path == CPU oracle on the whole read set (skipped with < 2 GPUs)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(root, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "multi-gpu ok" in out.stdout


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("k", [19, 30])
def test_total_kmer_counts_match_reference(golden, mods, case, k):
    """get_kmer_counts_reads on the device against the golden made by the reference's own function
    (better_consensus_unit_reconstruction.py:127-135, oracle/make_golden.py): distinct k-mers, total, largest count and
    the md5 of the sorted "kmer count" lines -- and against the loop restated here, entry by entry."""
    import hashlib
    import json
    import os
    from collections import Counter
    from centroflye_b200.better_consensus_unit_reconstruction import get_kmer_counts_reads
    _, _, NCRF_Report = mods
    rep = NCRF_Report(golden(case).report_path)
    got = get_kmer_counts_reads(rep, k=k)
    with open(os.path.join(golden(case).dir, f"kmer_counts_k{k}.json")) as f:
        gold = json.load(f)
    items = dict(got.items())
    assert len(items) == gold["n_distinct"] and sum(items.values()) == gold["n_total"] and max(items.values()) == gold["max_count"]
    lines = "".join(f"{kmer} {c}\n" for kmer, c in sorted(items.items()))
    assert hashlib.md5(lines.encode()).hexdigest() == gold["sorted_lines_md5"]
    want = Counter()
    for rec in rep.records.values():
        s = rec.r_al.replace("-", "")
        want.update(s[i:i + k] for i in range(len(s) - k + 1))
    assert items == dict(want)
    assert got["A" * k] == want.get("A" * k, 0) and got["not a kmer"] == 0


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("k", [19, 31])
def test_canonical_kmer_counts(golden, mods, case, k):
    """`jellyfish count -C` semantics (ext/tandemQUAST/scripts/select_kmers.py:131-133; the binary is an external
    dependency absent from the reference tree, so its published rule is restated): every k-mer occurrence counts for
    the smaller of the k-mer and its reverse complement (scripts/utils/bio.py:27-29 is the reference's RC)."""
    from collections import Counter
    from centroflye_b200.better_consensus_unit_reconstruction import get_canonical_kmer_counts, get_kmer_counts_reads
    _, _, NCRF_Report = mods
    rep = NCRF_Report(golden(case).report_path)
    comp = str.maketrans("ACGT", "TGCA")
    want = Counter()
    for rec in rep.records.values():
        s = rec.r_al.replace("-", "")
        for i in range(len(s) - k + 1):
            kmer = s[i:i + k]
            want[min(kmer, kmer.translate(comp)[::-1])] += 1
    got = dict(get_canonical_kmer_counts(rep, k=k).items())
    assert got == dict(want)
    if k == 19:  # merging strands conserves the total
        assert sum(got.values()) == sum(dict(get_kmer_counts_reads(rep, k=k).items()).values())
