"""a1/a2: the NCRF parser and the linear-scan unit segmentation against what the reference's
regex-based scripts/ncrf_parser.py produced (tests/golden/*/parsed.json.gz)."""
import hashlib

import pytest

from centroflye_b200.ncrf_parser import NCRF_Report, RC, motif_unit_columns
from conftest import golden_cases


@pytest.mark.parametrize("case", golden_cases())
def test_records_match_reference(golden, case):
    g = golden(case)
    rep = NCRF_Report(g.report_path)
    assert list(rep.records.keys()) == g.parsed["order"]
    assert sorted(rep.discarded_reads) == g.parsed["discarded"]
    for r_id, want in g.parsed["records"].items():
        rec = rep.records[r_id]
        got = dict(r_len=rec.r_len, r_al_len=rec.r_al_len, r_st=rec.r_st, r_en=rec.r_en, strand=rec.strand,
                   m_al_len=rec.m_al_len, al_score=rec.al_score,
                   r_al_md5=hashlib.md5(rec.r_al.encode()).hexdigest(),
                   m_al_md5=hashlib.md5(rec.m_al.encode()).hexdigest())
        assert got == want, r_id
    assert [sorted(x) for x in rep.classify(large_threshold=3000)] == g.parsed["classify_3000"]


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("n", [1, 2])
def test_segmentation_matches_reference(golden, case, n):
    g = golden(case)
    rep = NCRF_Report(g.report_path)
    for r_id, want in g.parsed["segments"][str(n)].items():
        mas = rep.records[r_id].get_motif_alignments(n=n)
        assert [[m.start, m.end] for m in mas] == want, r_id
        for m in mas:
            assert m.r_al == rep.records[r_id].r_al[m.start:m.end]
            assert m.m_al == rep.records[r_id].m_al[m.start:m.end]


# Known answers produced by running the reference, motif ACGTTGCA, 0.2*8 = 1.6
KAT = [
    ("ACGTTGCAACGTTGCA", "ACGTTGCAACGTTGCA", [0, 8, 16]),
    ("AACGTTGCAACGTTGCA", "AACGTTGCAACGTTGCA", [1, 9, 17]),            # 1-column prefix dropped
    ("CAACGTTGCAACGTTGCA", "CAACGTTGCAACGTTGCA", [0, 2, 10, 18]),       # 2-column prefix kept
    ("ACGTTGCAACGTTGCAA", "ACGTTGCAACGTTGCAA", [0, 8, 16]),            # 1-column suffix dropped
    ("ACGTTGCAACGTTGCAAC", "ACGTTGCAACGTTGCAAC", [0, 8, 16, 18]),       # 2-column suffix kept
    ("ACGTTGCATTACGTTGCA", "ACGTTGCA--ACGTTGCA", [0, 10, 18]),          # insertion joins preceding unit
    ("TTACGTTGCAACGTTGCA", "--ACGTTGCAACGTTGCA", [0, 2, 10, 18]),
    ("CGTTGCAACGTTGC", "CGTTGCAACGTTGC", []),
    ("ACGTTGCAACGTTGCA", "acgttgcaacgttgca", [0, 8, 16]),
]


@pytest.mark.parametrize("r_al,m_al,want", KAT)
def test_segmentation_known_answers(r_al, m_al, want):
    assert motif_unit_columns(m_al, len(r_al), "ACGTTGCA") == want


def test_segmentation_self_periodic_motif_is_leftmost_match():
    # motif ACAC is self-periodic: the leftmost match wins, exactly like the regex scan
    assert motif_unit_columns("ACACACACAC", 10, "ACAC") == [0, 4, 8, 10]
    assert motif_unit_columns("CACACACACA", 10, "ACAC") == [0, 1, 5, 9, 10]


def test_rc_passes_unknown_symbols():
    assert RC("ACGTacgt-N") == "N-acgtACGT"


def test_units_from_report_takes_reference_style_records(golden):
    """a report parsed by the REFERENCE's ncrf_parser.py (read_placer.py:9,106-114 keeps it when only the
    two hot-path modules are swapped) reaches the device path as plain objects with the reference's attributes -- no
    unit_columns method, no _cfk_source.  The segmentation must not depend on this repo's record class."""
    import types
    import numpy as np
    from centroflye_b200.ingest import batch_from_report, units_from_report
    rep = NCRF_Report(golden(golden_cases()[0]).report_path)
    duck = types.SimpleNamespace(records={
        r_id: types.SimpleNamespace(r_id=r.r_id, r_al=r.r_al, m_al=r.m_al, motif=r.motif, strand=r.strand,
                                    r_len=r.r_len, r_st=r.r_st, r_en=r.r_en)
        for r_id, r in rep.records.items()})
    for n in (1, 2):
        want = units_from_report(rep, batch_from_report(rep), n=n)
        got = units_from_report(duck, batch_from_report(duck), n=n)
        assert np.array_equal(got.read_unit_ptr, want.read_unit_ptr)
        assert np.array_equal(got.unit_off, want.unit_off) and np.array_equal(got.unit_len, want.unit_len)


def test_lazy_report_is_the_eager_report_once_looked_at(golden):
    """LazyNCRF_Report (what the command line builds) parses on first attribute access and is then indistinguishable
    from NCRF_Report; private _cfk attributes never trigger the parse."""
    from centroflye_b200.ncrf_parser import LazyNCRF_Report
    g = golden(golden_cases()[0])
    lazy, eager = LazyNCRF_Report(g.report_path), NCRF_Report(g.report_path)
    assert lazy._cfk_lazy_unparsed and getattr(lazy, "_cfk_cache", None) is None and lazy._cfk_lazy_unparsed
    assert isinstance(lazy, NCRF_Report)
    assert list(lazy.records.keys()) == list(eager.records.keys())        # first look: parsed now
    assert not lazy._cfk_lazy_unparsed
    for r_id, rec in eager.records.items():
        assert vars(lazy.records[r_id]) == vars(rec)
    assert lazy.read_lens == eager.read_lens and lazy.discarded_reads == eager.discarded_reads
    assert lazy.classify(large_threshold=3000) == eager.classify(large_threshold=3000)
    with pytest.raises(AttributeError):
        lazy.no_such_attribute
    with pytest.raises(FileNotFoundError):
        LazyNCRF_Report("/no/such/report.ncrf").records
