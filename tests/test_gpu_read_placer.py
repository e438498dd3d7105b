"""Row f rank 2: read_placer scoring on the device against read_positions.csv written by the UNMODIFIED reference
(scripts/read_placer.py:ReadPlacer.run, oracle/make_placer_golden.py) on the golden reports.  Placed reads must come
in the same greedy order with the same position and score; the trailing `r_id None` lines of unplaced reads are a set
in the reference (PYTHONHASHSEED-dependent order) and are compared as one."""
import argparse
import contextlib
import io
import json
import os

import numpy as np
import pytest

from conftest import golden_cases

pytestmark = pytest.mark.gpu


def _split(text):
    lines = text.splitlines()
    placed = [ln for ln in lines if not ln.endswith(" None")]
    return placed, sorted(ln for ln in lines if ln.endswith(" None"))


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("tag", ["default", "loose"])
def test_read_positions_match_reference(golden, tmp_path, case, tag):
    from centroflye_b200.encode import ints_to_kmers
    from centroflye_b200.read_placer import ReadPlacer
    g = golden(case)
    with open(os.path.join(g.dir, f"read_positions_{tag}.json")) as f:
        v = json.load(f)
    sel = np.load(os.path.join(g.dir, "p0.npz"))["selected"]
    kfn = tmp_path / "kmers.txt"
    kfn.write_text("".join(kmer + "\n" for kmer in ints_to_kmers(sel, v["k_cloud"])))
    params = argparse.Namespace(ncrf=g.report_path, genomic_kmers=str(kfn), k_cloud=v["k_cloud"], outdir=str(tmp_path / "out"),
                                n_motif=v["n_motif"], min_cloud_kmer_freq=v["min_cloud_kmer_freq"],
                                min_kmer_mult=v["min_kmer_mult"], min_unit=v["min_unit"], min_inters=v["min_inters"],
                                prefix_threshold=v["prefix_threshold"])
    with contextlib.redirect_stdout(io.StringIO()):
        placer = ReadPlacer(params)
        placer.run()
    got = open(os.path.join(params.outdir, "read_positions.csv")).read()
    want = open(os.path.join(g.dir, f"read_positions_{tag}.csv")).read()
    assert _split(got) == _split(want)
    placed = [ln.split() for ln in _split(got)[0]]
    assert placer.cloud_contig.read_positions == {p[0]: int(p[1]) for p in placed}


def test_placer_grows_its_score_tables(golden, tmp_path, monkeypatch):
    """Score tables that are too small report it; the call restores the contig and repeats itself with larger ones."""
    from centroflye_b200 import read_placer as rp
    from centroflye_b200.encode import ints_to_kmers
    g = golden("dxz1_small")
    v = json.load(open(os.path.join(g.dir, "read_positions_default.json")))
    sel = np.load(os.path.join(g.dir, "p0.npz"))["selected"]
    kfn = tmp_path / "kmers.txt"
    kfn.write_text("".join(kmer + "\n" for kmer in ints_to_kmers(sel, v["k_cloud"])))
    params = argparse.Namespace(ncrf=g.report_path, genomic_kmers=str(kfn), k_cloud=v["k_cloud"], outdir=str(tmp_path / "out"),
                                n_motif=1, min_cloud_kmer_freq=2, min_kmer_mult=2, min_unit=2, min_inters=10,
                                prefix_threshold=v["prefix_threshold"])
    once = rp.ReadPlacer._add_reads_once
    calls = []

    def tiny_first(self, reads, state, min_unit, min_inters, min_prop, grow):
        calls.append(grow)
        if len(calls) == 1:  # simulate the overflow path once, after the contig was touched
            self.cloud_contig.add_read(state, 0, state.r_ids[0], 7)
            return None
        return once(self, reads, state, min_unit, min_inters, min_prop, grow)
    monkeypatch.setattr(rp.ReadPlacer, "_add_reads_once", tiny_first)
    with contextlib.redirect_stdout(io.StringIO()):
        rp.ReadPlacer(params).run()
    assert calls[:2] == [1, 4]
    got = open(os.path.join(params.outdir, "read_positions.csv")).read()
    want = open(os.path.join(g.dir, "read_positions_default.csv")).read()
    assert _split(got) == _split(want)
