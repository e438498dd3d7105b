"""Host-side ingestion: alignment records -> the flat arrays the device consumes.

What the reference does per record and per stage, and what is shipped instead:

* stage A slides over ``r_al.replace('-', '')`` of every record
  (distance_based_kmer_recruitment.py:47-53)  ->  ``ReadBatch``: all gap-free
  read rows 2-bit packed back to back (every read starts on a 64-base boundary),
  plus ``read_off`` / ``read_len``;
* stage B slides over ``ma.r_al.upper().replace('-', '')`` of every unit
  returned by ``get_motif_alignments(n)`` (read_kmer_cloud.py:18-29)  ->
  ``UnitIndex``: for every unit its read, absolute start base and length in the
  same packed stream (units are sub-intervals of the read's gap-free row, so no
  second copy of the bases is made).

Buffers are numpy here; the engine copies them into pinned torch tensors.
"""
from dataclasses import dataclass, field

import numpy as np

from .encode import READ_ALIGN_BASES, ascii_to_codes, pack_codes
from .ncrf_parser import motif_unit_columns

_GAP = ord("-")


@dataclass
class ReadBatch:
    r_ids: list
    packed: np.ndarray      # uint32, 16 bases per word
    read_off: np.ndarray    # int64[R]   first base of read r in the packed base space
    read_len: np.ndarray    # int64[R]   gap-free length
    n_bases: int            # sum(read_len): the "read bases" of the metric

    @property
    def n_reads(self):
        return len(self.r_ids)

    def n_kmers(self, k):
        return int(np.maximum(self.read_len - k + 1, 0).sum())


@dataclass
class UnitIndex:
    read_unit_ptr: np.ndarray  # int64[R+1]  units of read r are [ptr[r], ptr[r+1])
    unit_off: np.ndarray       # int64[U]    absolute first base in the packed base space
    unit_len: np.ndarray       # int32[U]
    unit_read: np.ndarray      # int32[U]
    extra: dict = field(default_factory=dict)

    @property
    def n_units(self):
        return int(self.unit_off.size)

    def n_kmers(self, k):
        return int(np.maximum(self.unit_len.astype(np.int64) - k + 1, 0).sum())


def pack_reads(code_arrays, r_ids):
    """list of uint8 code arrays -> ReadBatch."""
    lens = np.array([a.size for a in code_arrays], dtype=np.int64)
    padded = (lens + READ_ALIGN_BASES - 1) // READ_ALIGN_BASES * READ_ALIGN_BASES
    off = np.zeros(len(code_arrays), dtype=np.int64)
    if len(code_arrays) > 1:
        off[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum())
    flat = np.zeros(total + READ_ALIGN_BASES, dtype=np.uint8)  # one spare 16-byte line for tail reads
    for a, o in zip(code_arrays, off):
        flat[o:o + a.size] = a
    return ReadBatch(r_ids=list(r_ids), packed=pack_codes(flat), read_off=off, read_len=lens,
                     n_bases=int(lens.sum()))


def gapfree_codes(r_al):
    row = np.frombuffer(r_al.encode("latin-1"), dtype=np.uint8)
    return ascii_to_codes(row[row != _GAP])


def batch_from_report(report):
    """All records of an ``NCRF_Report`` in dict order (the order every reference loop uses)."""
    r_ids = list(report.records.keys())
    return pack_reads([gapfree_codes(report.records[r].r_al) for r in r_ids], r_ids)


def units_from_boundaries(batch, boundaries):
    """boundaries[r] = gap-free offsets b_0 <= b_1 <= ... of read r's units (may be empty)."""
    counts = np.array([max(len(b) - 1, 0) for b in boundaries], dtype=np.int64)
    ptr = np.zeros(len(boundaries) + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    U = int(ptr[-1])
    unit_off = np.empty(U, dtype=np.int64)
    unit_len = np.empty(U, dtype=np.int32)
    unit_read = np.empty(U, dtype=np.int32)
    for r, b in enumerate(boundaries):
        if counts[r] == 0:
            continue
        b = np.asarray(b, dtype=np.int64)
        lo, hi = ptr[r], ptr[r + 1]
        unit_off[lo:hi] = batch.read_off[r] + b[:-1]
        unit_len[lo:hi] = b[1:] - b[:-1]
        unit_read[lo:hi] = r
    return UnitIndex(read_unit_ptr=ptr, unit_off=unit_off, unit_len=unit_len, unit_read=unit_read)


def units_from_report(report, batch, n=1):
    """Unit boundaries of every record, converted from alignment columns to gap-free offsets."""
    boundaries = []
    for r_id in batch.r_ids:
        rec = report.records[r_id]
        # any record with the reference's attributes will do (a report parsed by the reference's own ncrf_parser.py,
        # read_placer.py:9,106-114): the segmentation is a free function, not a method of this repo's record class
        coords = motif_unit_columns(rec.m_al, len(rec.r_al), rec.motif, n=n)
        if not coords:
            boundaries.append(np.empty(0, dtype=np.int64))
            continue
        row = np.frombuffer(rec.r_al.encode("latin-1"), dtype=np.uint8)
        before = np.concatenate([[0], np.cumsum(row != _GAP)])
        boundaries.append(before[np.asarray(coords, dtype=np.int64)])
    return units_from_boundaries(batch, boundaries)


def batch_from_synth(reads, motif_len, min_record_len=5000):
    """Synthetic reads straight to device form, applying the parser's record rule
    (longest alignment per id is moot: ids are unique) and min_record_len (ncrf_parser.py:91-93)."""
    kept, codes, bounds = [], [], []
    for rd in reads:
        if rd.r_al_len < min_record_len:
            continue
        bases, b = rd.direct_units(motif_len)
        kept.append(rd.r_id)
        codes.append(bases)
        bounds.append(b)
    batch = pack_reads(codes, kept)
    return batch, units_from_boundaries(batch, bounds)


def native_ingest(report_fn, n=1, min_record_len=5000, threads=0):
    """NCRF report file -> (ReadBatch, UnitIndex, fields) in one native pass (libcfk.so, csrc/ncrf_ingest.cpp).

    Equal, array for array, to ``batch_from_report(NCRF_Report(fn))`` + ``units_from_report(report, batch, n)``
    (tests/test_ncrf_native.py), without a Python object or string per record.  ``fields`` is an int64 array
    [R, 8]: r_len, r_al_len, r_st, r_en, strand (+1/-1), m_al_len, score, alignment columns — the scalar
    attributes of the reference's NCRF_Record (scripts/ncrf_parser.py:13-26) after the strand flip (:96-100).
    Raises ValueError on what makes the Python path raise (dangling line, malformed record, non-ACGT symbol)."""
    import ctypes
    import os
    from . import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.cfk_ncrf_open(os.fsencode(report_fn), int(min_record_len), int(n), int(threads), ctypes.byref(ctx))
    if rc != 0:
        msg = lib.cfk_ncrf_last_error().decode(errors="replace")
        if "cannot open" in msg:
            raise FileNotFoundError(msg)
        raise ValueError(msg)
    try:
        R, U = int(lib.cfk_ncrf_n_records(ctx)), int(lib.cfk_ncrf_n_units(ctx))
        packed = np.empty(int(lib.cfk_ncrf_n_words(ctx)), dtype=np.uint32)
        read_off, read_len = np.empty(R, dtype=np.int64), np.empty(R, dtype=np.int64)
        ptr = np.empty(R + 1, dtype=np.int64)
        unit_off, unit_len = np.empty(U, dtype=np.int64), np.empty(U, dtype=np.int32)
        unit_read = np.empty(U, dtype=np.int32)
        ids = np.empty(max(int(lib.cfk_ncrf_ids_bytes(ctx)), 1), dtype=np.uint8)
        fields = np.empty((R, 8), dtype=np.int64)
        rc = lib.cfk_ncrf_export(ctx, packed.ctypes.data, read_off.ctypes.data, read_len.ctypes.data, ptr.ctypes.data,
                                 unit_off.ctypes.data, unit_len.ctypes.data, unit_read.ctypes.data, ids.ctypes.data,
                                 fields.ctypes.data)
        if rc != 0:
            raise ValueError(lib.cfk_ncrf_last_error().decode(errors="replace"))
        n_bases = int(lib.cfk_ncrf_n_bases(ctx))
    finally:
        lib.cfk_ncrf_close(ctx)
    r_ids = ids.tobytes()[: int(ids.size) if R else 0].decode("utf-8", errors="replace").split("\n")[:R]
    batch = ReadBatch(r_ids=r_ids, packed=packed, read_off=read_off, read_len=read_len, n_bases=n_bases)
    return batch, UnitIndex(read_unit_ptr=ptr, unit_off=unit_off, unit_len=unit_len, unit_read=unit_read), fields
