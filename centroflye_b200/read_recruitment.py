"""The read recruitment pre-filter of the reference (``scripts/read_recruitment/rr.cpp``) on the
device: same command line (``rr unit.fasta reads.fasta[.gz] output.fasta edit_distance_threshold``), same output file.

rr.cpp:73-90 aligns the unit and its reverse complement against every read with edlib in HW mode (the whole unit
against any infix of the read, k = threshold) and copies the read to the output when either alignment exists.  Here the
two alignments of every read are bit-parallel scans on the GPU (csrc/rr_filter.cu, cfk_rr_filter); the host parses the
FASTA / FASTQ (plain or gzip, like kseq) and writes ``>name\\nsequence\\n`` for the kept reads in input order.
"""
import gzip
import sys

import numpy as np

from . import _lib
from ._lib import CfkError

_COMPLEMENT = {"A": "T", "T": "A", "G": "C", "C": "G"}
SEGMENT = 8192  # read bases per device work item (a multiple of 16: segments start on 16-byte boundaries)


def reverse_complement(unit):
    """rr.cpp:11-39,58-63; a unit symbol outside ACGT trips the reference's assert."""
    try:
        return "".join(_COMPLEMENT[c] for c in reversed(unit))
    except KeyError as e:
        raise AssertionError(f"complement(): unexpected symbol {e.args[0]!r} in the unit") from None


def read_sequences(path):
    """(name, sequence) of every FASTA / FASTQ record, plain or gzip -- what kseq_read yields: the name is the header up
    to the first whitespace, sequence lines are joined until a line starting with '>', '@' or '+'; after '+' as many
    quality characters as the sequence has bases are skipped."""
    with open(path, "rb") as f:
        magic = f.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    with opener(path, "rt") as f:
        line = f.readline()
        while line:
            if line[:1] not in (">", "@"):
                line = f.readline()
                continue
            name = (line[1:].split() or [""])[0]
            chunks = []
            line = f.readline()
            while line and line[:1] not in (">", "@", "+"):
                chunks.append(line.strip())
                line = f.readline()
            seq = "".join(chunks)
            if line[:1] == "+":
                left = len(seq)
                line = f.readline()
                while line and left > 0:
                    left -= len(line.rstrip("\r\n"))
                    line = f.readline()
            yield name, seq


def build_masks(unit):
    """-> (peq uint64[2, S, NW], sym_of uint8[256], NW) for cfk_rr_filter."""
    lib = _lib.load()
    m = len(unit)
    nw = int(lib.cfk_rr_words(m))
    if nw < 0:
        raise CfkError(f"read recruitment: the unit has {m} bases; 1..3328 are supported")
    n_sym = int(lib.cfk_rr_max_symbols())
    symbols = sorted(set((unit + reverse_complement(unit)).encode("latin-1")))  # both strands share one symbol table
    if len(symbols) > n_sym - 1:
        raise CfkError(f"read recruitment: the unit uses {len(symbols)} distinct symbols; at most {n_sym - 1} are supported")
    sym_of = np.zeros(256, dtype=np.uint8)
    for slot, b in enumerate(symbols, start=1):
        sym_of[b] = slot
    peq = np.zeros((2, n_sym, nw), dtype=np.uint64)
    for strand, seq in enumerate((unit, reverse_complement(unit))):
        codes = sym_of[np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)]
        pos = np.arange(m)
        for slot in range(1, len(symbols) + 1):
            hit = pos[codes == slot]
            np.bitwise_or.at(peq[strand, slot], hit >> 6, np.uint64(1) << (hit & 63).astype(np.uint64))
    return peq, sym_of, nw


def recruit(unit, seqs, threshold, exact=False, engine=None):
    """keep[r] (bool) for every sequence, and with exact=True the infix edit distances int32[R, 2] (unit, reverse
    complement) of cfk_rr_filter."""
    from .engine import default_engine
    eng = engine or default_engine()
    t = eng.torch
    peq, sym_of, _ = build_masks(unit)
    R = len(seqs)
    keep = np.zeros(R, dtype=bool)
    dist = np.zeros((R, 2), dtype=np.int32) if exact else None
    d_peq, d_sym = eng._to_dev(peq.reshape(-1).view(np.int64)), eng._to_dev(sym_of)
    start, budget = 0, 1 << 30  # device batches of about 1 GiB of read bytes
    while start < R:
        end, total = start, 0
        while end < R and (end == start or total + len(seqs[end]) <= budget):
            total += (len(seqs[end]) + 15) & ~15
            end += 1
        lens = np.array([len(s) for s in seqs[start:end]], dtype=np.int64)
        offs = np.zeros(end - start, dtype=np.int64)
        np.cumsum(((lens + 15) & ~15)[:-1], out=offs[1:])
        text = np.zeros(int(offs[-1] + ((lens[-1] + 15) & ~15)) + 16, dtype=np.uint8)
        for o, s in zip(offs.tolist(), seqs[start:end]):
            text[o:o + len(s)] = np.frombuffer(s.encode("latin-1"), dtype=np.uint8)
        # One thread walks one segment.  A hit within `threshold` edits spans at most |unit| + threshold read bases, so a
        # long read is cut into segments of SEGMENT bases that overlap by that much: the decision per read (any segment
        # hits) is unchanged, the longest read no longer sets the kernel's duration.  Exact distances need whole reads.
        seg_read = np.arange(end - start, dtype=np.int64)
        seg_off, seg_len = offs, lens
        if not exact and threshold >= 0:
            overlap = len(unit) + int(threshold)
            n_seg = np.maximum(1, -(-np.maximum(lens - overlap, 1) // SEGMENT))
            seg_read = np.repeat(seg_read, n_seg)
            first = np.cumsum(n_seg) - n_seg
            j = np.arange(int(n_seg.sum()), dtype=np.int64) - np.repeat(first, n_seg)
            seg_off = offs[seg_read] + j * SEGMENT
            seg_len = np.minimum(SEGMENT + overlap, lens[seg_read] - j * SEGMENT)
        order = np.argsort(-seg_len, kind="stable").astype(np.int32)
        n = int(seg_read.size)
        d_keep = eng._zeros(n, t.uint8)
        d_dist = eng._empty(2 * n, t.int32) if exact else None
        d_text, d_offs, d_lens, d_order = eng._to_dev(text), eng._to_dev(seg_off), eng._to_dev(seg_len), eng._to_dev(order)
        with eng._stage("rr_filter"):
            _lib.call("cfk_rr_filter", eng._p(d_text), eng._p(d_offs), eng._p(d_lens), eng._p(d_order), n, eng._p(d_peq),
                      eng._p(d_sym), len(unit), int(threshold), int(bool(exact)), eng._p(d_dist), eng._p(d_keep), eng._stream())
        hit = np.zeros(end - start, dtype=bool)
        np.logical_or.at(hit, seg_read, d_keep[:n].cpu().numpy().astype(bool))
        keep[start:end] = hit
        if exact:
            dist[start:end] = d_dist[: 2 * n].cpu().numpy().reshape(n, 2)
        start = end
    return (keep, dist) if exact else keep


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    if len(argv) != 4 or argv[0] == "-h":
        print("Usage: ./rr unit.fasta reads.fasta.gz output.fasta edit_distance_threshold", end="")
        return 0
    unit_fn, read_fn, output_fn = argv[:3]
    try:
        threshold = int(argv[3])
    except ValueError:
        threshold = 0  # std::atoi
    unit = next(read_sequences(unit_fn))[1]
    records = list(read_sequences(read_fn))
    keep = recruit(unit, [seq for _, seq in records], threshold) if records else np.zeros(0, dtype=bool)
    with open(output_fn, "w") as f:
        f.write("".join(f">{name}\n{seq}\n" for (name, seq), k in zip(records, keep.tolist()) if k))
    return 0


if __name__ == "__main__":
    sys.exit(main())
