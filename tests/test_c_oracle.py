"""Pins oracle/c/cfk_oracle.c (the large-input checker and bench.py's CPU baseline) to the reference's
outputs under tests/golden/, through the same flat arrays the device consumes."""
import numpy as np
import pytest

from centroflye_b200.engine import band_to_int
from centroflye_b200.ingest import batch_from_report, units_from_report
from centroflye_b200.ncrf_parser import NCRF_Report
from conftest import all_case_points
from oracle import c_oracle


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("case,pi", all_case_points())
def test_c_oracle_matches_reference(golden, case, pi, threads):
    g = golden(case)
    meta, arr = g.point(pi)
    p = meta["params"]
    k = p["k"]
    rep = NCRF_Report(g.report_path)
    batch = batch_from_report(rep)
    units = units_from_report(rep, batch, n=1)
    lo, hi = band_to_int(p["bottom"] * p["coverage"] * p["kmer_survival_rate"],
                         p["top"] * p["coverage"] * p["kmer_survival_rate"])
    n_reads = batch.n_reads
    lo_r, hi_r, _ = slice(p["min_nreads"], p["max_nreads"]).indices(n_reads)
    hi_r = max(hi_r, lo_r)
    out = c_oracle.recruit(batch, units, k, lo, hi, p["max_nonuniq"], p["min_distance"], p["max_distance"],
                           p["min_coverage"], threads=threads, unit_lo=int(units.read_unit_ptr[lo_r]),
                           unit_hi=int(units.read_unit_ptr[hi_r]))
    assert np.array_equal(out["all_keys"], arr["all_keys"])
    assert np.array_equal(out["all_counts"], arr["all_counts"])
    assert np.array_equal(out["rare"], arr["rare"])
    assert np.array_equal(out["unit_ptr"], arr["clouds_unit_ptr"])
    assert np.array_equal(out["rare"][out["ids"]], arr["clouds_kmers"])
    assert np.array_equal(np.diff(units.read_unit_ptr), arr["clouds_units_per_read"])
    assert out["n_increments"] == meta["n_increments"]
    e = out["edges"]
    order = np.lexsort((e[:, 1], e[:, 0], e[:, 2]))
    e = e[order]
    assert np.array_equal(e[:, 2].astype(np.int32), arr["edge_d"])
    assert np.array_equal(out["rare"][e[:, 0]], arr["edge_a"])
    assert np.array_equal(out["rare"][e[:, 1]], arr["edge_b"])
    assert np.array_equal(e[:, 3], arr["edge_cnt"])
    assert np.array_equal(out["rare"][out["selected"]], arr["selected"])


def test_sampled_baseline_extrapolates(golden):
    g = golden("rand311")
    rep = NCRF_Report(g.report_path)
    batch = batch_from_report(rep)
    units = units_from_report(rep, batch, n=1)
    params = dict(k=19, max_nonuniq=3, min_d=1, max_d=150, min_coverage=4)
    res = c_oracle.timed_sample(batch, units, params, (4, 11), bounded_s=5.0, threads=2)
    assert res["value"] > 0 and res["kind"] == "port" and res["cores"] == 2
    assert res["sources_done"] == res["sources"]  # small input: the budget is never hit
