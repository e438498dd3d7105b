"""The synthetic generator: text route (NCRF report -> parser -> segmentation) and direct route
(truth columns -> device arrays) must describe the same reads and units."""
import numpy as np

from centroflye_b200 import synth
from centroflye_b200.encode import pack_codes, unpack_codes
from centroflye_b200.ingest import batch_from_report, batch_from_synth, units_from_report
from centroflye_b200.ncrf_parser import NCRF_Report


def _dataset(tmp_path, unit_len=211, mult=90, err=0.06):
    unit = synth.random_unit(unit_len, 1)
    genome, a0, alen = synth.simulate_genome(unit, mult, 0.02, 2, flank_len=4000)
    reads = synth.simulate_reads(genome, a0, alen, unit, 8, err, 3, median_len=7000, sigma=0.4, min_len=3000,
                                 max_len=20000)
    path = tmp_path / "synth.ncrf"
    synth.write_ncrf_report(path, reads, unit)
    return unit, reads, str(path)


def test_text_route_equals_direct_route(tmp_path):
    unit, reads, path = _dataset(tmp_path)
    rep = NCRF_Report(path)
    b1 = batch_from_report(rep)
    u1 = units_from_report(rep, b1, n=1)
    b2, u2 = batch_from_synth(reads, len(unit))
    assert b1.r_ids == b2.r_ids and len(b1.r_ids) > 5
    assert any(rd.strand == "-" for rd in reads) and any(rd.strand == "+" for rd in reads)
    assert np.array_equal(b1.packed, b2.packed)
    assert np.array_equal(b1.read_off, b2.read_off) and np.array_equal(b1.read_len, b2.read_len)
    assert np.array_equal(u1.read_unit_ptr, u2.read_unit_ptr)
    assert np.array_equal(u1.unit_off, u2.unit_off) and np.array_equal(u1.unit_len, u2.unit_len)
    assert np.array_equal(u1.unit_read, u2.unit_read)
    assert b1.n_bases == sum(len(r.r_al.replace("-", "")) for r in rep.records.values())
    # units tile a contiguous stretch of each read
    for r in range(b1.n_reads):
        lo, hi = u1.read_unit_ptr[r], u1.read_unit_ptr[r + 1]
        if hi > lo:
            ends = u1.unit_off[lo:hi] + u1.unit_len[lo:hi]
            assert np.array_equal(ends[:-1], u1.unit_off[lo + 1:hi])
            assert ends[-1] <= b1.read_off[r] + b1.read_len[r]


def test_short_alignments_are_dropped_like_the_parser(tmp_path):
    unit, reads, path = _dataset(tmp_path)
    rep = NCRF_Report(path)
    short = [rd.r_id for rd in reads if rd.r_al_len < 5000]
    assert short and all(r not in rep.records for r in short)


def test_pack_roundtrip():
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 4, size=1003, dtype=np.uint8)
    assert np.array_equal(unpack_codes(pack_codes(codes), 1003), codes)


def test_generator_is_deterministic(tmp_path):
    a = _dataset(tmp_path)[1]
    b = _dataset(tmp_path)[1]
    assert all(np.array_equal(x.r_row, y.r_row) and np.array_equal(x.m_row, y.m_row) for x, y in zip(a, b))
