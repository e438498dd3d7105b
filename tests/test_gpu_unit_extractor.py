"""Row f rank 4: the k-mer analysis of scripts/unit_extractor.py:23-105 with the position grouping on the device, against
outputs of the reference's own functions (oracle/make_unit_extractor_golden.py compiled them out of the reference file) on
reads of the golden reports: repetitive k-mers with their position lists in the reference's dict order, per-k-mer gap
lists, the sorted union, periods / bins, the hook k-mer and the splits."""
import hashlib
import json
import os

import pytest

from conftest import golden_cases

pytestmark = pytest.mark.gpu


def _digest(obj):
    return hashlib.md5(json.dumps(obj, sort_keys=False, separators=(",", ":")).encode()).hexdigest()


@pytest.mark.parametrize("case", golden_cases())
def test_unit_extractor_functions_match_reference(golden, case):
    from centroflye_b200 import unit_extractor as ue
    from centroflye_b200.ncrf_parser import NCRF_Report
    g = golden(case)
    rep = NCRF_Report(g.report_path)
    with open(os.path.join(g.dir, "unit_extractor.json")) as f:
        want = json.load(f)
    for w in want:
        seq = rep.records[w["r_id"]].r_al.replace("-", "").upper()
        assert len(seq) == w["seq_len"]
        rep_kmers = ue.get_repetitive_kmers(seq, w["k"])
        assert len(rep_kmers) == w["n_rep_kmers"]
        assert _digest(list(rep_kmers.items())) == w["rep_kmers_md5"]          # same keys, same order, same lists
        conv, union_conv = ue.get_convolution(rep_kmers)
        assert _digest(list(conv.items())) == w["conv_md5"]
        assert len(union_conv) == w["n_union_conv"] and _digest(union_conv) == w["union_conv_md5"]
        conv2, union2 = ue.get_convolution(dict(rep_kmers))                    # a plain dict takes the host arithmetic
        assert conv2 == conv and union2 == union_conv
        periods, bin_convs, bin_left, bin_right = ue.get_period_info(union_conv, 10)
        assert list(periods)[:20] == w["periods"] and list(bin_convs)[:20] == w["bin_convs"]
        assert (bin_left, bin_right) == (w["bin_left"], w["bin_right"])
        hook = ue.get_hook_kmer(conv, bin_left, bin_right) if union_conv else None
        assert hook == w["hook"]
        splits = ue.split_by_hook(seq, hook) if hook else {}
        assert list(splits.keys()) == w["split_ids"] and _digest(list(splits.items())) == w["splits_md5"]


def test_unit_extractor_edge_cases():
    from centroflye_b200 import unit_extractor as ue
    assert ue.get_repetitive_kmers("ACG", 5) == {}
    assert ue.get_repetitive_kmers("ACGTACGA", 4) == {}                        # nothing repeats
    r = ue.get_repetitive_kmers("AAAAAA", 3)
    assert r == {"AAA": [0, 1, 2, 3]}
    assert ue.get_convolution(r) == ({"AAA": [1, 1, 1]}, [1, 1, 1])
    assert ue.get_period_info([], 10) == ([], [], None, None)
    assert ue.split_by_hook("ACGTTACGTTACG", "ACG") == {"split_0_5": "ACGTT", "split_5_10": "ACGTT"}
    with pytest.raises(ValueError):
        ue.get_repetitive_kmers("ACGNACGT", 3)
