"""Row (f1): the native NCRF ingestion (libcfk.so, csrc/ncrf_ingest.cpp) against the Python host path, which
tests/test_parser_golden.py pins to the reference's regex parser: same records in the same order, same packed
reads, same unit table -- on the golden reports and on hand-made reports hitting every parser rule
(scripts/ncrf_parser.py:61-118, :28-59)."""
import numpy as np
import pytest

from centroflye_b200.ingest import batch_from_report, native_ingest, units_from_report
from centroflye_b200.ncrf_parser import NCRF_Report, RC
from conftest import golden_cases


def _same(report_fn, n, min_record_len=5000, threads=0):
    rep = NCRF_Report(report_fn, min_record_len=min_record_len)
    want_b = batch_from_report(rep)
    want_u = units_from_report(rep, want_b, n=n)
    batch, units, fields = native_ingest(report_fn, n=n, min_record_len=min_record_len, threads=threads)
    assert batch.r_ids == want_b.r_ids
    assert batch.n_bases == want_b.n_bases
    assert np.array_equal(batch.read_off, want_b.read_off) and np.array_equal(batch.read_len, want_b.read_len)
    assert np.array_equal(batch.packed, want_b.packed)
    assert np.array_equal(units.read_unit_ptr, want_u.read_unit_ptr)
    assert np.array_equal(units.unit_off, want_u.unit_off) and np.array_equal(units.unit_len, want_u.unit_len)
    assert np.array_equal(units.unit_read, want_u.unit_read)
    for i, r_id in enumerate(batch.r_ids):
        rec = rep.records[r_id]
        assert fields[i].tolist() == [rec.r_len, rec.r_al_len, rec.r_st, rec.r_en, -1 if rec.strand == "-" else 1,
                                      rec.m_al_len, rec.al_score, len(rec.r_al)], r_id
    return batch, units


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("n", [1, 2])
def test_native_ingest_equals_python_path_on_goldens(golden, case, n):
    batch, units = _same(golden(case).report_path, n, threads=1 if n == 1 else 0)
    assert batch.n_reads > 0 and units.n_units > 0


MOTIF = "ACGTTGCA"


def _record(r_id, r_al, m_al, strand="+", r_len=None, r_al_len=None, st=0, score=7, sep=" "):
    r_len = r_len if r_len is not None else len(r_al.replace("-", "")) + 10
    r_al_len = r_al_len if r_al_len is not None else len(r_al.replace("-", ""))
    en = st + r_al_len
    if strand == "-":  # the file holds the reverse-complemented rows (ncrf_parser.py:96-100 flips them back)
        r_al, m_al = RC(r_al), RC(m_al)
    return (f"{r_id}{sep}{r_len} {r_al_len}bp {st}-{en} {r_al}\n"
            f"{MOTIF}{strand} {len(m_al.replace('-', ''))}bp score={score} {m_al}\n")


def _write(tmp_path, text):
    p = tmp_path / "report.ncrf"
    p.write_text(text)
    return str(p)


def test_native_ingest_parser_rules(tmp_path):
    two = MOTIF * 2
    text = "# header comment\n\n   \n"
    text += _record("r_plain", two, two)
    text += _record("r_prefix_dropped", "A" + two, "A" + two)
    text += _record("r_prefix_kept", "CA" + two, "CA" + two)
    text += _record("r_suffix_kept", two + "AC", two + "AC")
    text += _record("r_insertion", "ACGTTGCATTACGTTGCA", "ACGTTGCA--ACGTTGCA")
    text += _record("r_deletion", "ACG-TGCAACGTTGCA", "ACGTTGCAACGTTGCA")
    text += _record("r_leading_gap", "TT" + two, "--" + two)
    text += _record("r_nomatch", "CGTTGCAACGTTGC", "CGTTGCAACGTTGC")
    text += _record("r_lower_motif_row", two, two.lower())
    text += _record("r_minus", "CA" + two + "G", "CA" + two + "G", strand="-", st=3)
    text += "# a comment between records\n"
    text += _record("r_dup", two, two)                                   # kept first ...
    text += _record("r_short", two, two, r_al_len=3)                     # below min_record_len: discarded
    text += _record("r_dup", two + MOTIF, two + MOTIF)                   # ... replaced by the longer alignment, same slot
    text += _record("r_dup", MOTIF, MOTIF)                               # shorter: ignored
    text += _record("r_tab", two, two, sep="\t")                         # \s+ after the id may be a tab (regex backtracking)
    text += "\n"
    fn = _write(tmp_path, text)
    for n in (1, 2):
        batch, units = _same(fn, n, min_record_len=8)
    assert "r_short" not in batch.r_ids and batch.r_ids.index("r_dup") == 10 and "r_tab" in batch.r_ids
    _same(fn, 1, min_record_len=0)  # now r_short is kept


def test_native_ingest_empty_report(tmp_path):
    fn = _write(tmp_path, "# nothing here\n\n")
    batch, units, fields = native_ingest(fn)
    assert batch.n_reads == 0 and units.n_units == 0 and batch.n_bases == 0 and fields.shape == (0, 8)
    assert units.read_unit_ptr.tolist() == [0]


def test_native_ingest_reads_what_cannot_be_mapped(tmp_path):
    """The report is mapped read-only; a FIFO (the reference's open() reads one just as well) and a zero-byte file take
    the read-to-the-end path and give the same arrays."""
    import os
    import threading
    two = MOTIF * 2
    text = _record("r_a", two + "ACG", two + "ACG") + "# comment\n" + _record("r_b", "T" + two, "T" + two, strand="-")
    want, want_units, want_fields = native_ingest(_write(tmp_path, text), min_record_len=0)
    fifo = str(tmp_path / "report.fifo")
    os.mkfifo(fifo)
    feeder = threading.Thread(target=lambda: open(fifo, "w").write(text))
    feeder.start()
    got, got_units, got_fields = native_ingest(fifo, min_record_len=0)
    feeder.join()
    assert got.r_ids == want.r_ids and np.array_equal(got.packed, want.packed) and np.array_equal(got.read_len, want.read_len)
    assert np.array_equal(got_units.unit_off, want_units.unit_off) and np.array_equal(got_fields, want_fields)
    empty = tmp_path / "empty.ncrf"
    empty.write_bytes(b"")
    batch, units, _ = native_ingest(str(empty))
    assert batch.n_reads == 0 and units.n_units == 0


def test_native_ingest_errors(tmp_path):
    two = MOTIF * 2
    with pytest.raises(ValueError, match="non-ACGT"):
        native_ingest(_write(tmp_path, _record("r", two + "N", two + "A")), min_record_len=0)
    with pytest.raises(ValueError, match="dangling"):
        native_ingest(_write(tmp_path, _record("r", two, two) + "r2 30 16bp 0-16 ACGT\n"), min_record_len=0)
    with pytest.raises(ValueError, match="malformed"):
        native_ingest(_write(tmp_path, "r 30 16 0-16 ACGT\n" + f"{MOTIF}+ 16bp score=1 ACGT\n"), min_record_len=0)
    with pytest.raises(FileNotFoundError):
        native_ingest(str(tmp_path / "missing.ncrf"))


@pytest.mark.parametrize("seed", range(6))
def test_native_ingest_random_reports(tmp_path, seed):
    """Random small reports: gaps in both rows, both strands, lower-case motif rows, repeated read ids with longer and
    shorter alignments, comment and blank lines, units cut at the ends -- native ingestion == Python host path."""
    rng = np.random.default_rng(seed)
    motif = "".join(rng.choice(list("ACGT"), size=int(rng.integers(5, 40))))
    lines = ["# random report", ""]
    ids = [f"read_{i}" for i in range(12)]
    for rec in range(30):
        r_id = ids[int(rng.integers(0, len(ids)))]
        n_units = int(rng.integers(0, 6))
        cut_l, cut_r = int(rng.integers(0, len(motif))), int(rng.integers(0, len(motif)))
        truth = (motif * (n_units + 2))[cut_l: len(motif) * (n_units + 2) - cut_r]
        r_row, m_row = [], []
        for ch in truth:  # truth alignment with substitutions, insertions and deletions
            x = rng.random()
            if x < 0.05:
                r_row.append(str(rng.choice(list("ACGT")))); m_row.append(ch)
            elif x < 0.09:
                r_row.append("-"); m_row.append(ch)
            elif x < 0.13:
                r_row.append(str(rng.choice(list("ACGT")))); m_row.append("-")
                r_row.append(ch); m_row.append(ch)
            else:
                r_row.append(ch); m_row.append(ch)
        r_al, m_al = "".join(r_row), "".join(m_row)
        if not r_al.replace("-", ""):
            continue
        if rng.random() < 0.3:
            m_al = m_al.lower()
        strand = "-" if rng.random() < 0.5 else "+"
        n_bases = len(r_al.replace("-", ""))
        if strand == "-":
            r_al, m_al = RC(r_al), RC(m_al)
        lines.append(f"{r_id} {n_bases + 20} {n_bases}bp 5-{5 + n_bases} {r_al}")
        lines.append(f"{motif}{strand} {len(m_al.replace('-', ''))}bp score={int(rng.integers(0, 999))} {m_al}")
        if rng.random() < 0.2:
            lines.append("# interleaved comment")
        if rng.random() < 0.2:
            lines.append("   ")
    fn = _write(tmp_path, "\n".join(lines) + "\n")
    for n in (1, 2, 3):
        _same(fn, n, min_record_len=int(rng.integers(0, 60)), threads=int(rng.integers(0, 3)))
