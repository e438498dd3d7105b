"""NCRF report input boundary of the recruitment path.

Mirrors the public surface of the reference's ``scripts/ncrf_parser.py``
(``NCRF_Report``, ``NCRF_Report.NCRF_Record``, ``get_motif_alignments``) with
the same record-selection and strand rules, but the unit segmentation is a
linear scan instead of a 2055-capture-group regular expression:

* ``NCRF_Report.__init__``      <- scripts/ncrf_parser.py:61-118
* ``get_motif_alignments``      <- scripts/ncrf_parser.py:28-59
* ``RC``                        <- scripts/utils/bio.py:27-29

Segmentation restatement.  The reference matches ``b0([-]*)b1([-]*)...`` (the
motif, each base followed by any number of '-') non-overlapping and leftmost
over ``m_al.upper()``.  A match can only start on a non-gap column, consumes
exactly len(motif)*n non-gap symbols that spell the motif, and greedily eats
the gap columns behind its last base.  So the matches are exactly the
non-overlapping leftmost occurrences of ``motif*n`` in the gap-free motif row,
mapped back to alignment columns: start = column of the first matched symbol,
end = column of the next non-gap symbol (or the row length).  ``str.find`` on
the gap-free row does that in linear time.
"""
import re
from collections import defaultdict, namedtuple

import numpy as np

_RC_TABLE = str.maketrans("ATGCatgc-", "TACGtacg-")

MotifAlignment = namedtuple("MotifAlignment", ["r_id", "start", "end", "r_al", "m_al"])

_FIRST_LINE = re.compile(r"^([^ ]+)\s+(\d+)\s+(\d+)bp\s+(\d+)-(\d+)\s+(.+)$")
_SECOND_LINE = re.compile(r"^([^+-]+)([+-])\s+(\d+)bp\s+score=(\d+)\s+(.+)$")

_GAP = ord("-")


def RC(s):
    """Reverse complement; symbols outside ``ATGCatgc-`` pass through (utils/bio.py:27-29)."""
    return s.translate(_RC_TABLE)[::-1]


def motif_unit_columns(m_al, r_al_len, motif, n=1, overlapped=False):
    """Column boundaries of the units of one alignment (scripts/ncrf_parser.py:34-52).

    Returns a list ``coords`` such that unit j spans alignment columns
    ``[coords[j], coords[j+1])``; empty list when the motif never matches.
    """
    pattern = motif * n
    row = np.frombuffer(m_al.upper().encode("latin-1"), dtype=np.uint8)
    cols = np.flatnonzero(row != _GAP)  # column of every non-gap symbol
    flat = row[cols].tobytes().decode("latin-1")
    plen = len(pattern)
    if plen == 0:
        return []
    starts = []
    q = flat.find(pattern)
    while q != -1:
        starts.append(q)
        q = flat.find(pattern, q + 1 if overlapped else q + plen)
    if not starts:
        return []
    coords = [int(cols[q]) for q in starts]
    last_end = starts[-1] + plen
    coords.append(int(cols[last_end]) if last_end < cols.size else len(m_al))
    # partial first / last unit kept only when longer than 20 % of the motif (:49-52)
    if coords[0] > len(motif) * 0.2:
        coords.insert(0, 0)
    if coords[-1] < r_al_len - len(motif) * 0.2:
        coords.append(r_al_len)
    return coords


class NCRF_Report:
    class NCRF_Record:
        def __init__(self, r_id, r_len, r_al_len, r_st, r_en, r_al,
                     motif, strand, m_al_len, al_score, m_al):
            self.r_id = r_id
            self.r_len = int(r_len)
            self.r_al_len = int(r_al_len)
            self.r_st = int(r_st)
            self.r_en = int(r_en)
            self.r_al = r_al
            self.motif = motif
            self.strand = strand
            self.m_al_len = int(m_al_len)
            self.al_score = int(al_score)
            self.m_al = m_al

        def unit_columns(self, n=1, overlapped=False):
            return motif_unit_columns(self.m_al, len(self.r_al), self.motif, n=n, overlapped=overlapped)

        def get_motif_alignments(self, n=1, overlapped=False):
            coords = self.unit_columns(n=n, overlapped=overlapped)
            return [MotifAlignment(r_id=self.r_id, start=st, end=en,
                                   r_al=self.r_al[st:en], m_al=self.m_al[st:en])
                    for st, en in zip(coords[:-1], coords[1:])]

    def __init__(self, report_fn, min_record_len=5000):
        self.records = {}
        self._cfk_source = (report_fn, min_record_len)  # lets the device path ingest the file natively (ingest.native_ingest)
        self.positions_all_alignments = defaultdict(list)
        self.read_lens = {}
        with open(report_fn, "r") as f:
            lines = [ln.strip() for ln in f]
        lines = [ln for ln in lines if ln and ln[0] != "#"]
        seen = set()
        for p in range(0, len(lines), 2):
            fst, snd = lines[p:p + 2]  # ValueError on a dangling line, like the reference
            r_id, r_len, r_al_len, r_st, r_en, r_al = _FIRST_LINE.search(fst).groups()
            motif, strand, m_al_len, al_score, m_al = _SECOND_LINE.search(snd).groups()
            r_len, r_al_len, r_st, r_en = int(r_len), int(r_al_len), int(r_st), int(r_en)
            seen.add(r_id)
            self.positions_all_alignments[r_id].append((r_st, r_en, strand))
            self.read_lens[r_id] = r_len
            held = self.records.get(r_id)
            if held is not None and held.r_al_len >= r_al_len:
                continue
            if r_al_len < min_record_len:
                continue
            if strand == "-":
                # alignment is flipped to the motif's forward strand; the strand
                # field keeps saying '-' (scripts/ncrf_parser.py:96-100)
                r_st, r_en = r_len - r_en, r_len - r_st
                r_al, m_al = RC(r_al), RC(m_al)
            self.records[r_id] = self.NCRF_Record(
                r_id=r_id, r_len=r_len, r_al_len=r_al_len, r_st=r_st, r_en=r_en, r_al=r_al,
                motif=motif, strand=strand, m_al_len=int(m_al_len), al_score=int(al_score), m_al=m_al)
        for r_id in self.positions_all_alignments:
            self.positions_all_alignments[r_id].sort()
        self.discarded_reads = [r_id for r_id in seen if r_id not in self.records]

    def classify(self, large_threshold, small_threshold=1000):
        """Prefix / internal / suffix read classes (scripts/ncrf_parser.py:120-145)."""
        prefix_reads, suffix_reads, internal_reads = [], [], []
        for r_id, record in self.records.items():
            r_len = self.read_lens[r_id]
            spans = self.positions_all_alignments[r_id]
            if record.strand == "+":
                left_pos, right_pos = spans[0][0], spans[-1][1]
            else:
                left_pos, right_pos = r_len - spans[-1][1], r_len - spans[0][0]
            if (left_pos > large_threshold and right_pos > r_len - small_threshold
                    and right_pos == record.r_en):
                prefix_reads.append(r_id)
            elif (right_pos < r_len - large_threshold and left_pos < small_threshold
                    and left_pos == record.r_st):
                suffix_reads.append(r_id)
            else:
                internal_reads.append(r_id)
        return prefix_reads, internal_reads, suffix_reads

    def get_efficiency(self):
        """Share of each read's aligned span kept by the longest record (:147-164)."""
        efficiency, total, used = {}, 0, 0
        for r_id, spans in self.positions_all_alignments.items():
            span_len = sum(en - st + 1 for st, en, _ in spans)
            total += span_len
            record = self.records.get(r_id)
            if record is None:
                efficiency[r_id] = 0
            else:
                kept = record.r_en - record.r_st + 1
                used += kept
                efficiency[r_id] = kept / span_len
        return efficiency, used / total

    def get_motif_alignments(self, n=1):
        """r_id -> unit alignments of its record (scripts/ncrf_parser.py:170-174)."""
        return {r_id: record.get_motif_alignments(n=n) for r_id, record in self.records.items()}


class LazyNCRF_Report(NCRF_Report):
    """An NCRF_Report that parses its file into Python records only when somebody looks at them.

    The command line of distance_based_kmer_recruitment.py hands the report straight to the device path, which ingests
    the FILE natively (csrc/ncrf_ingest.cpp) and never touches ``records``: building 10^4 record objects with 80 kB
    strings each (0.5 s for a 320 MB report) would be the slowest step of the run.  Any access to an attribute of the
    eager class (``records``, ``read_lens``, ``classify`` ...) runs the eager constructor first, so callers see the same
    object either way; while it is unparsed nobody can have edited a record, so the device path may trust the file."""

    def __init__(self, report_fn, min_record_len=5000):
        self.__dict__["_cfk_source"] = (report_fn, min_record_len)
        self.__dict__["_cfk_lazy_unparsed"] = True

    def __getattr__(self, name):  # only called for attributes that are not there (yet)
        if name.startswith("_cfk") or not self.__dict__.get("_cfk_lazy_unparsed", False):
            raise AttributeError(name)
        self.__dict__["_cfk_lazy_unparsed"] = False
        NCRF_Report.__init__(self, *self.__dict__["_cfk_source"])
        return getattr(self, name)
