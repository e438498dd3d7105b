"""Drop-in for the reference's ``scripts/read_kmer_cloud.py`` (same names, same call shapes).

  ReadKMerCloud(kmers, r_id), .fromNCRF_record(...)          read_kmer_cloud.py:9-31
  get_reads_kmer_clouds(ncrf_report, n, k, genomic_kmers)    read_kmer_cloud.py:34-40
  filter_reads_kmer_clouds(kmer_clouds, min_mult, max_mult)  read_kmer_cloud.py:43-54

The clouds live on the device as a CSR (read -> units -> sorted k-mer ids); the Python
objects the reference's callers expect (``dict[r_id -> ReadKMerCloud]`` whose ``.kmers`` is a
``list[set[str]]``) are views that materialise strings only when someone looks at them
(read_placer.py:47-49, cloud_contig.py:29-33 do; the recruitment script never does).
"""
import math
from dataclasses import dataclass

import numpy as np

from .encode import check_k, ints_to_kmers, kmers_to_ints
from .ingest import batch_from_report, units_from_report


@dataclass
class CloudState:
    """Device-resident clouds of one call + the host facts needed to interpret them."""
    engine: object
    csr: object             # engine.CloudCSR
    index: object           # engine.KmerIndex (sorted keys; rank = id)
    unit_last: object       # device int32[U]
    read_unit_ptr: np.ndarray
    r_ids: list
    k: int
    _host: tuple = None

    def host(self):
        if self._host is None:
            self._host = (self.csr.unit_ptr.cpu().numpy(), self.csr.ids.cpu().numpy().view(np.uint32),
                          self.index.sorted_keys.cpu().numpy().view(np.uint64))
        return self._host

    def read_sets(self, r):
        unit_ptr, ids, keys = self.host()
        u0, u1 = int(self.read_unit_ptr[r]), int(self.read_unit_ptr[r + 1])
        lo, hi = int(unit_ptr[u0]), int(unit_ptr[u1])
        strs = ints_to_kmers(keys[ids[lo:hi]], self.k)
        return [set(strs[int(unit_ptr[u]) - lo:int(unit_ptr[u + 1]) - lo]) for u in range(u0, u1)]

    def host_unit_sizes(self):
        """int64[U]: cloud size of every unit (host copy of the CSR pointer, made once)."""
        return np.diff(self.host()[0][: self.csr.n_units + 1]) if self.csr.n_units else np.zeros(0, dtype=np.int64)

    def with_csr(self, csr):
        return CloudState(self.engine, csr, self.index, self.unit_last, self.read_unit_ptr, self.r_ids, self.k)


class ReadKMerCloud:
    def __init__(self, kmers, r_id):
        self.r_id = r_id
        self._kmers = kmers
        self._state = self._all_state = self._owner = None
        self._read = 0
        self._all = None

    @classmethod
    def _view(cls, r_id, state, read_index, owner):
        self = cls.__new__(cls)
        self.r_id = r_id
        self._kmers = None
        self._state = state
        self._all_state = state  # all_kmers is a construction-time snapshot (read_kmer_cloud.py:13-15)
        self._read = read_index
        self._all = None
        self._owner = owner
        return self

    @property
    def kmers(self):
        if self._kmers is None:
            self._kmers = self._state.read_sets(self._read)
            if self._owner is not None:
                self._owner._host_touched = True  # the caller may now mutate the sets behind the device's back
        return self._kmers

    @kmers.setter
    def kmers(self, value):
        self._kmers = value
        if self._owner is not None:
            self._owner._host_touched = True

    @property
    def all_kmers(self):
        if self._all is None:
            sets = self._all_state.read_sets(self._read) if self._all_state is not None else self._kmers
            self._all = [kmer for unit in sets for kmer in unit]
        return self._all

    @classmethod
    def fromNCRF_record(cls, ncrf_record, n, k, genomic_kmers):
        class _One:
            records = {ncrf_record.r_id: ncrf_record}
        return get_reads_kmer_clouds(_One(), n=n, k=k, genomic_kmers=genomic_kmers)[ncrf_record.r_id]


class CloudDict(dict):
    """``dict[r_id -> ReadKMerCloud]`` in record order + the device state behind it."""
    _state = None
    _host_touched = False

    def device_state(self):
        """CloudState if the device copy is still authoritative, else None."""
        if self._state is None or self._host_touched or list(self.keys()) != self._state.r_ids:
            return None
        return self._state


# ---- per-report caches (parse once, upload once) ---------------------------------------------
def _unparsed(report):
    """A LazyNCRF_Report nobody has looked into yet: its file is the only copy of the records."""
    return bool(getattr(report, "_cfk_lazy_unparsed", False))


def _report_cache(report):
    cache = getattr(report, "_cfk_cache", None)
    n_records = -1 if _unparsed(report) else len(report.records)
    if cache is None or (n_records >= 0 and cache.get("n_records") not in (-1, n_records)):
        cache = {"n_records": n_records, "units": {}, "dev_reads": {}, "dev_units": {}}
        try:
            report._cfk_cache = cache
        except AttributeError:
            pass
    return cache


def _native(report, n):
    """(ReadBatch, UnitIndex) of the report's file through the native ingestion, or None when the report object was
    not built from a file or no longer matches it (records added / removed by the caller)."""
    src = getattr(report, "_cfk_source", None)
    if src is None:
        return None
    from .ingest import native_ingest
    try:
        batch, units, _ = native_ingest(src[0], n=n, min_record_len=src[1])
    except (OSError, ValueError):
        return None  # the Python path below raises the reference-shaped error, if any
    if _unparsed(report):
        return batch, units  # no Python record exists that could disagree with the file
    if batch.r_ids != list(report.records.keys()):
        return None
    # a caller may have edited records in place since the file was parsed: the file is only trusted while every
    # record still has the gap-free length the file gave it (C-level scans; the bases themselves are not compared)
    for r_id, n_bases in zip(batch.r_ids, batch.read_len.tolist()):
        r_al = report.records[r_id].r_al
        if len(r_al) - r_al.count("-") != n_bases:
            return None
    return batch, units


def report_batch(report):
    cache = _report_cache(report)
    if "batch" not in cache:
        got = _native(report, 1)
        if got is not None:
            cache["batch"], cache["units"][1] = got
        else:
            cache["batch"] = batch_from_report(report)
    return cache["batch"]


def report_units(report, n):
    cache = _report_cache(report)
    if n not in cache["units"]:
        got = _native(report, n) if getattr(report, "_cfk_source", None) is not None else None
        if got is not None and np.array_equal(got[0].read_off, report_batch(report).read_off):
            cache["units"][n] = got[1]
        else:
            cache["units"][n] = units_from_report(report, report_batch(report), n=n)
    return cache["units"][n]


def report_device_reads(report, engine, k):
    cache = _report_cache(report)
    key = (id(engine), k)
    if key not in cache["dev_reads"]:
        cache["dev_reads"][key] = engine.upload_reads(report_batch(report), k)
    return cache["dev_reads"][key]


def report_device_units(report, engine, n, k):
    cache = _report_cache(report)
    key = (id(engine), n, k)
    if key not in cache["dev_units"]:
        cache["dev_units"][key] = engine.upload_units(report_units(report, n), k)
    return cache["dev_units"][key]


def index_for_kmers(engine, genomic_kmers, k):
    """set[str] (or a RareKmerSet carrying its device index) -> KmerIndex."""
    idx = getattr(genomic_kmers, "_cfk_index", None)
    if idx is not None and getattr(genomic_kmers, "_cfk_k", None) == k and idx[0] is engine:
        return idx[1]
    acgt = set("ACGT")
    # k-mers of another length or alphabet can never equal a read k-mer: dropping them changes nothing
    usable = [kmer for kmer in genomic_kmers if isinstance(kmer, str) and len(kmer) == k and not (set(kmer) - acgt)]
    return engine.index_from_host_keys(kmers_to_ints(usable, k))


# ---- the reference's functions ---------------------------------------------------------------
def get_reads_kmer_clouds(ncrf_report, n, k, genomic_kmers=None):
    from .engine import default_engine
    if genomic_kmers is None:
        raise TypeError("argument of type 'NoneType' is not iterable")  # read_kmer_cloud.py:28 with the default
    k = check_k(k)
    engine = default_engine()
    batch = report_batch(ncrf_report)
    units = report_units(ncrf_report, n)
    reads = report_device_reads(ncrf_report, engine, k)
    dev_units = report_device_units(ncrf_report, engine, n, k)
    index = index_for_kmers(engine, genomic_kmers, k)
    csr = engine.build_clouds(reads, dev_units, k, index)
    state = CloudState(engine, csr, index, dev_units.unit_last, units.read_unit_ptr, batch.r_ids, k)
    out = CloudDict()
    out._state = state
    for r, r_id in enumerate(batch.r_ids):
        out[r_id] = ReadKMerCloud._view(r_id, state, r, out)
    return out


def state_from_sets(kmer_clouds, kmers=None):
    """Any ``dict[r_id -> object with .kmers: list[iterable[str]]]`` -> CloudState on the device.

    Converting the caller's strings is host work; every count is still done by the device.
    ``kmers`` fixes the id universe (its sorted order); a cloud k-mer outside it is a KeyError,
    as at distance_based_kmer_recruitment.py:96."""
    from .engine import CloudCSR, default_engine
    from .ingest import UnitIndex
    engine = default_engine()
    r_ids = list(kmer_clouds.keys())
    flat, sizes, per_read = [], [], []
    for r_id in r_ids:
        units = kmer_clouds[r_id].kmers
        per_read.append(len(units))
        for unit in units:
            sizes.append(len(unit))
            flat.extend(unit)
    universe = sorted(set(kmers)) if kmers is not None else sorted(set(flat))
    k = len(universe[0]) if universe else (len(flat[0]) if flat else 1)
    keys = np.sort(kmers_to_ints(universe, k)) if universe else np.empty(0, dtype=np.uint64)
    vals = kmers_to_ints(flat, k) if flat else np.empty(0, dtype=np.uint64)
    ids = np.searchsorted(keys, vals)
    if vals.size:
        probe = keys[np.minimum(ids, max(keys.size - 1, 0))] if keys.size else np.full(vals.size, ~np.uint64(0))
        bad = np.flatnonzero(probe != vals)
        if bad.size:
            raise KeyError(flat[int(bad[0])])
    unit_ptr = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(np.asarray(sizes, dtype=np.int64), out=unit_ptr[1:])
    ids = ids.astype(np.uint32)
    for lo, hi in zip(unit_ptr[:-1], unit_ptr[1:]):
        ids[lo:hi].sort()
    read_unit_ptr = np.zeros(len(r_ids) + 1, dtype=np.int64)
    np.cumsum(np.asarray(per_read, dtype=np.int64), out=read_unit_ptr[1:])
    U = len(sizes)
    units = UnitIndex(read_unit_ptr=read_unit_ptr, unit_off=np.zeros(U, dtype=np.int64),
                      unit_len=np.zeros(U, dtype=np.int32),
                      unit_read=np.repeat(np.arange(len(r_ids), dtype=np.int32), per_read))
    dev_units = engine.upload_units(units, k)
    index = engine.index_from_host_keys(keys)
    csr = CloudCSR(unit_ptr=engine._to_dev(unit_ptr), ids=engine._to_dev(ids.view(np.int32)), n_units=U,
                   n_entries=int(ids.size))
    return CloudState(engine, csr, index, dev_units.unit_last, read_unit_ptr, r_ids, k)


def filter_reads_kmer_clouds(kmer_clouds, min_mult=2, max_mult=math.inf):
    state = kmer_clouds.device_state() if isinstance(kmer_clouds, CloudDict) else None
    if state is not None:  # untouched device clouds: filter there, re-point the views
        new_state = state.with_csr(state.engine.filter_clouds(state.csr, state.index.n, min_mult, max_mult))
        kmer_clouds._state = new_state
        for r, r_id in enumerate(state.r_ids):
            view = kmer_clouds[r_id]
            view._state, view._read, view._kmers = new_state, r, None
        kmer_clouds._host_touched = False
        return kmer_clouds
    # host-held sets (a foreign dict, or views somebody already looked at): count on the device,
    # then write the surviving sets back in place exactly like read_kmer_cloud.py:49-53
    state = state_from_sets(kmer_clouds)
    new_state = state.with_csr(state.engine.filter_clouds(state.csr, state.index.n, min_mult, max_mult))
    for r, r_id in enumerate(state.r_ids):
        for pos, kept in enumerate(new_state.read_sets(r)):
            kmer_clouds[r_id].kmers[pos] = kept
    return kmer_clouds


def get_all_kmers(kmer_clouds):
    """read_kmer_cloud.py:57-63 AS WRITTEN: the loop reads ``kmer_clouds.all_kmers`` -- an attribute of the dict, not
    of the cloud it iterates over -- so a non-empty argument raises AttributeError and an empty one returns [].
    Nothing in the reference calls it; it is here so that ``from read_kmer_cloud import *`` binds the same names with
    the same behaviour."""
    merged = []
    for _ in kmer_clouds:
        merged += kmer_clouds.all_kmers
    merged.sort()
    assert len(set(merged)) == len(merged)
    return merged
