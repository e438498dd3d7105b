"""Host-side halves of the row-f items (no GPU): the FASTA / FASTQ reader and match masks of the rr drop-in, and the
host-side functions of the unit_extractor drop-in against the goldens made from the reference's functions."""
import gzip
import hashlib
import json
import os
from collections import defaultdict

import numpy as np
import pytest

from conftest import golden_cases


def test_read_sequences_like_kseq(tmp_path):
    from centroflye_b200.read_recruitment import read_sequences
    fa = tmp_path / "a.fasta"
    fa.write_text(">r1 first read\nACGT\nACG\n\n>r2\nTTTT\n>empty\n>r3\tx\nA\n")
    assert list(read_sequences(str(fa))) == [("r1", "ACGTACG"), ("r2", "TTTT"), ("empty", ""), ("r3", "A")]
    fq = tmp_path / "a.fastq"
    fq.write_text("@q1 c\nACGT\n+\n@@@@\n@q2\nAC\nGT\n+q2\n@I\nII\n")  # quality lines may start with '@'
    assert list(read_sequences(str(fq))) == [("q1", "ACGT"), ("q2", "ACGT")]
    gz = tmp_path / "a.fasta.gz"
    with gzip.open(gz, "wt") as f:
        f.write(">z\nGATTACA\n")
    assert list(read_sequences(str(gz))) == [("z", "GATTACA")]


def test_rr_masks_and_reverse_complement():
    from centroflye_b200.read_recruitment import build_masks, reverse_complement
    assert reverse_complement("AACGT") == "ACGTT"
    unit = "ACGTT" * 30  # 150 bases: three 64-bit words, template size 4
    peq, sym_of, nw = build_masks(unit)
    assert nw == 4 and peq.shape[0] == 2 and peq.shape[2] == 4
    for strand, seq in enumerate((unit, reverse_complement(unit))):
        for i, ch in enumerate(seq):
            slot = int(sym_of[ord(ch)])
            assert slot > 0
            for s in range(1, 5):
                bit = (int(peq[strand, s, i >> 6]) >> (i & 63)) & 1
                assert bit == (1 if s == slot else 0)
        assert not peq[strand, :, 3].any() and all(int(x) >> 22 == 0 for x in peq[strand, :, 2])  # nothing behind base 149
    assert int(sym_of[ord("N")]) == 0 and not peq[:, 0].any()  # a symbol the unit lacks matches nothing


def _digest(obj):
    return hashlib.md5(json.dumps(obj, sort_keys=False, separators=(",", ":")).encode()).hexdigest()


@pytest.mark.parametrize("case", golden_cases())
def test_unit_extractor_host_functions(golden, case):
    """get_convolution (plain-dict path), get_period_info, get_hook_kmer, split_by_hook against the goldens of the
    reference's functions; the repetitive k-mers come from the loop of unit_extractor.py:23-31 restated here."""
    from centroflye_b200 import unit_extractor as ue
    from centroflye_b200.ncrf_parser import NCRF_Report
    g = golden(case)
    rep = NCRF_Report(g.report_path)
    with open(os.path.join(g.dir, "unit_extractor.json")) as f:
        want = json.load(f)
    for w in want:
        seq = rep.records[w["r_id"]].r_al.replace("-", "").upper()
        pos = defaultdict(list)
        for i in range(len(seq) - w["k"] + 1):
            pos[seq[i:i + w["k"]]].append(i)
        rep_kmers = {kmer: p for kmer, p in pos.items() if len(p) > 1}
        assert _digest(list(rep_kmers.items())) == w["rep_kmers_md5"]
        conv, union_conv = ue.get_convolution(rep_kmers)
        assert _digest(list(conv.items())) == w["conv_md5"] and _digest(union_conv) == w["union_conv_md5"]
        periods, bin_convs, bin_left, bin_right = ue.get_period_info(union_conv, 10)
        assert list(periods)[:20] == w["periods"] and list(bin_convs)[:20] == w["bin_convs"]
        assert (bin_left, bin_right) == (w["bin_left"], w["bin_right"])
        hook = ue.get_hook_kmer(conv, bin_left, bin_right) if union_conv else None
        assert hook == w["hook"]
        splits = ue.split_by_hook(seq, hook) if hook else {}
        assert list(splits.keys()) == w["split_ids"] and _digest(list(splits.items())) == w["splits_md5"]
