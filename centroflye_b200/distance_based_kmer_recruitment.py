"""Drop-in for the reference's ``scripts/distance_based_kmer_recruitment.py``.

Same command line (``parse_args``, dbkr.py:15-36), same function names and call shapes
(:39, :66, :85, :131, :152, :174), same two output files — computed by libcfk.so on a B200.

Where the reference returns containers that cannot exist at scale (a dict with one entry per
distinct k-mer, a table with one counter per (distance, k-mer, k-mer)), the objects returned
here implement the same reading interface over device / numpy arrays:

  get_kmer_freqs_from_ncrf_report -> KmerFreqs     mapping kmer -> n_reads (missing -> 0, as defaultdict)
  get_rare_kmers                  -> RareKmerSet   a real ``set[str]`` that also carries the device index
  get_kmer_dist_map               -> (DistCounts, KmerRanks)  ``dist_cnt[d][i][j]``, ``kmer_index[kmer]``
  filter_dist_tuples              -> (set[int], EdgeList)     tuples (dist, i, j, freq), sorted

Ids are canonical: ``kmer_index[kmer]`` is the rank of the k-mer in sorted order, and edges are
sorted by (dist, i, j); the reference's orders depend on PYTHONHASHSEED.
"""
import argparse
import os
import sys
from collections import defaultdict
from collections.abc import Mapping, Sequence

import numpy as np

from .encode import check_k, ints_to_kmers, kmers_to_ints
from .engine import band_to_int, default_engine, to_host_u32, to_host_u64, U32_MAX
from .ncrf_parser import LazyNCRF_Report, NCRF_Report
from .read_kmer_cloud import (CloudDict, get_reads_kmer_clouds, report_device_reads, state_from_sets)
from .utils.os_utils import smart_makedirs


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--ncrf', help='NCRF report on reads', required=True)
    parser.add_argument('--coverage', help='Average coverage of the dataset', type=int, required=True)
    parser.add_argument('--min-coverage', help='minCov threshold', type=int, default=4)
    parser.add_argument('--outdir', help='Output directory', required=True)
    parser.add_argument('-k', type=int, default=19)
    parser.add_argument('--min-nreads', type=int, default=0)
    parser.add_argument('--max-nreads', type=int, default=sys.maxsize)
    parser.add_argument('--min-distance', type=int, default=1)
    parser.add_argument('--max-distance', type=int, default=150)
    parser.add_argument('--bottom', type=float, default=0.9)
    parser.add_argument('--top', type=float, default=3.)
    parser.add_argument('--kmer-survival-rate', type=float, default=0.34)
    parser.add_argument('--max-nonuniq', type=int, default=3)
    parser.add_argument('--verbose', action='store_true', default=True)
    return parser.parse_args(argv)


# ---- result containers ----------------------------------------------------------------------
class KmerFreqs(Mapping):
    """kmer -> number of records containing it (keys sorted); reads like the reference's defaultdict."""

    def __init__(self, keys_u64, counts, k, table=None, engine=None):
        order = np.argsort(keys_u64)
        self.keys_u64 = keys_u64[order]
        self.counts = counts[order]
        self.k = k
        self._table = table
        self._engine = engine

    def _find(self, kmer):
        if not isinstance(kmer, str) or len(kmer) != self.k or set(kmer) - set("ACGT"):
            return -1
        v = kmers_to_ints([kmer], self.k)[0]
        i = int(np.searchsorted(self.keys_u64, v))
        return i if i < self.keys_u64.size and self.keys_u64[i] == v else -1

    def __getitem__(self, kmer):
        i = self._find(kmer)
        return int(self.counts[i]) if i >= 0 else 0

    def __contains__(self, kmer):
        return self._find(kmer) >= 0

    def __len__(self):
        return int(self.keys_u64.size)

    def __iter__(self):
        return iter(ints_to_kmers(self.keys_u64, self.k))

    def items(self):
        return zip(ints_to_kmers(self.keys_u64, self.k), (int(c) for c in self.counts))

    def values(self):
        return (int(c) for c in self.counts)


class RareKmerSet(set):
    """``set[str]`` of rare k-mers that remembers its sorted device copy (ids = ranks).  Any in-place change of the
    set drops the device copy: the next consumer rebuilds it from the strings.

    main() never looks at the strings (every consumer on its path takes the device copy), so there the ~10^6 Python
    strings are not made up front: ``_cfk_pending`` holds the sorted keys and the first Python-level look at the set
    (len, iteration, membership, any mutator, ``materialize()``) fills them in.  get_rare_kmers() called by anybody
    else returns the filled set, as the reference does."""
    _cfk_index = None
    _cfk_k = None
    _cfk_pending = None

    def materialize(self):
        if self._cfk_pending is not None:
            keys, self._cfk_pending = self._cfk_pending, None
            set.update(self, ints_to_kmers(keys, self._cfk_k))
        return self

    def __len__(self):
        return set.__len__(self.materialize())

    def __iter__(self):
        return set.__iter__(self.materialize())

    def __contains__(self, kmer):
        return set.__contains__(self.materialize(), kmer)


def _dropping_index(name):
    method = getattr(set, name)

    def mutate(self, *args, **kwargs):
        self.materialize()
        self._cfk_index = None
        return method(self, *args, **kwargs)
    mutate.__name__ = name
    return mutate


for _name in ("add", "discard", "remove", "pop", "clear", "update", "difference_update", "intersection_update",
              "symmetric_difference_update", "__ior__", "__iand__", "__isub__", "__ixor__"):
    setattr(RareKmerSet, _name, _dropping_index(_name))


def _filled_first(name):
    method = getattr(set, name)

    def read(self, *args, **kwargs):
        return method(self.materialize(), *args, **kwargs)
    read.__name__ = name
    return read


# the read-only methods of set look at the C-level table directly: fill it first.  (A plain set built FROM a pending
# RareKmerSet -- set(rare), other | rare -- cannot be intercepted; only main() ever holds a pending one.)
for _name in ("copy", "union", "intersection", "difference", "symmetric_difference", "issubset", "issuperset",
              "isdisjoint", "__or__", "__and__", "__sub__", "__xor__", "__ror__", "__rand__", "__rsub__", "__rxor__",
              "__eq__", "__ne__", "__le__", "__lt__", "__ge__", "__gt__", "__reduce__", "__repr__"):
    setattr(RareKmerSet, _name, _filled_first(_name))
RareKmerSet.__hash__ = None  # as for set: defining __eq__ must not make instances hashable


class KmerRanks(Mapping):
    """kmer -> id, id = rank in sorted order (the reference's ``kmer_index``, dbkr.py:103)."""

    def __init__(self, sorted_keys_u64, k):
        self.keys_u64 = sorted_keys_u64
        self.k = k

    def __getitem__(self, kmer):
        v = kmers_to_ints([kmer], self.k)[0] if isinstance(kmer, str) and len(kmer) == self.k else None
        if v is not None:
            i = int(np.searchsorted(self.keys_u64, v))
            if i < self.keys_u64.size and self.keys_u64[i] == v:
                return i
        raise KeyError(kmer)

    def __len__(self):
        return int(self.keys_u64.size)

    def __iter__(self):
        return iter(ints_to_kmers(self.keys_u64, self.k))

    def items(self):
        return zip(ints_to_kmers(self.keys_u64, self.k), range(self.keys_u64.size))

    def kmer_of(self, ids):
        return ints_to_kmers(self.keys_u64[np.asarray(ids, dtype=np.int64)], self.k)


class EdgeList(Sequence):
    """Edges as the reference's tuples ``(dist, i, j, freq)``, sorted by (dist, i, j), numpy-backed.

    Built from four columns, or (``from_rows``) from the device's own rows uint32[n][4] = (i, j, dist, freq): then the
    int64 columns ``dist`` / ``i`` / ``j`` / ``freq`` are made only if somebody reads them -- the native writer takes
    the rows as they are."""

    def __init__(self, dist, i, j, freq, presorted=False):
        if not presorted:
            order = np.lexsort((j, i, dist))
            dist, i, j, freq = dist[order], i[order], j[order], freq[order]
        self.rows = None
        self._cols = tuple(np.ascontiguousarray(x, dtype=np.int64) for x in (dist, i, j, freq))

    @classmethod
    def from_rows(cls, rows, presorted=False):
        rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1, 4)
        if not presorted:
            rows = rows[np.lexsort((rows[:, 1], rows[:, 0], rows[:, 2]))]
        self = cls.__new__(cls)
        self.rows, self._cols = rows, None
        return self

    def _columns(self):
        if self._cols is None:
            self._cols = tuple(self.rows[:, c].astype(np.int64) for c in (2, 0, 1, 3))
        return self._cols

    dist = property(lambda self: self._columns()[0])
    i = property(lambda self: self._columns()[1])
    j = property(lambda self: self._columns()[2])
    freq = property(lambda self: self._columns()[3])

    def __len__(self):
        return int(self.rows.shape[0]) if self._cols is None else int(self._cols[0].size)

    def __getitem__(self, n):
        if isinstance(n, slice):
            return [self[t] for t in range(*n.indices(len(self)))]
        return (int(self.dist[n]), int(self.i[n]), int(self.j[n]), int(self.freq[n]))

    def __iter__(self):
        return zip(self.dist.tolist(), self.i.tolist(), self.j.tolist(), self.freq.tolist())


class DistCounts(Mapping):
    """The reference's ``dist_cnt`` (dbkr.py:108-127) as a lazy handle.

    ``filter_dist_tuples`` consumes it on the device without ever expanding it.  Reading it like
    the reference's dict (``dist_cnt[d][i][j]``, ``.items()``) expands every non-zero counter on
    the device (min_coverage = 1) and is meant for small inputs and tests."""

    def __init__(self, state, unit_lo, unit_hi, min_d, max_d):
        self.state = state
        self.unit_lo, self.unit_hi = unit_lo, unit_hi
        self.min_d, self.max_d = min_d, max_d
        self._tables = None
        self.n_increments = None
        self._occ = None

    def occurrences(self):
        if self._occ is None:
            st = self.state
            self._occ = st.engine.build_occurrences(st.csr, st.index.n, self.unit_lo, self.unit_hi, st.unit_last)
        return self._occ

    def run(self, min_coverage, rel_threshold):
        st = self.state
        return st.engine.dist_edges(st.csr, st.unit_last, st.index.n, self.min_d, self.max_d, min_coverage,
                                    rel_threshold, unit_lo=self.unit_lo, unit_hi=self.unit_hi,
                                    occurrences=self.occurrences() if st.index.n and st.csr.n_entries else None)

    def _expand(self):
        if self._tables is None:
            res = self.run(1, 0.0)  # every non-zero counter passes
            e = to_host_u32(res.edges).reshape(-1, 4)
            n = self.state.index.n
            tables = {d: [defaultdict(int) for _ in range(n)] for d in range(self.min_d, self.max_d + 1)}
            for a, b, d, c in e.tolist():
                tables[d][a][b] = c
            self._tables = tables
            self.n_increments = res.n_increments
        return self._tables

    def __getitem__(self, dist):
        return self._expand()[dist]

    def __iter__(self):
        return iter(range(self.min_d, self.max_d + 1))

    def __len__(self):
        return max(self.max_d - self.min_d + 1, 0)


# ---- the reference's functions ---------------------------------------------------------------
def _count_table(reads_ncrf_report, k):
    engine = default_engine()
    reads = report_device_reads(reads_ncrf_report, engine, k)
    return engine, engine.count_docfreq(reads, k)


def get_kmer_freqs_from_ncrf_report(reads_ncrf_report, k, verbose, max_nonuniq):
    k = check_k(k)
    engine, table = _count_table(reads_ncrf_report, k)
    if max_nonuniq < 0:  # dbkr.py:57-62: every k-mer fails `non_unique_freqs[kmer] <= max_nonuniq` and is deleted
        keys, nreads = engine._empty(0, engine.torch.int64)[:0], engine._empty(0, engine.torch.int32)[:0]
    else:
        keys, nreads, _ = engine.table_select(table, 0, U32_MAX, max_nonuniq, with_counts=True)
    if verbose:
        print(len(reads_ncrf_report.records), len(reads_ncrf_report.records))
    return KmerFreqs(to_host_u64(keys), to_host_u32(nreads), k, table=table, engine=engine)


def get_rare_kmers(reads_ncrf_report, k, bottom, top, coverage, kmer_survival_rate, max_nonuniq, verbose,
                   _strings=True):
    k = check_k(k)
    engine = default_engine()
    reads = report_device_reads(reads_ncrf_report, engine, k)
    left = bottom*coverage*kmer_survival_rate   # float64, this order (dbkr.py:74-75)
    right = top*coverage*kmer_survival_rate
    lo, hi = band_to_int(left, right)
    rare_keys = engine.rare_kmers(reads, k, lo, hi, max_nonuniq)  # counting and band in one device pass
    index = engine.build_index(rare_keys)
    rare = RareKmerSet()
    rare._cfk_index = (engine, index)
    rare._cfk_k = k
    rare._cfk_pending = to_host_u64(index.sorted_keys)
    if _strings:
        rare.materialize()
    if verbose:
        print(f'# rare kmers: {index.n}')
    return rare


def get_kmer_dist_map(reads_kmer_clouds, kmers, min_n, max_n, min_d, max_d, verbose):
    if min_d < 0:
        raise ValueError("min_d < 0 is not meaningful for the unit-distance loop (dbkr.py:121)")
    state = reads_kmer_clouds.device_state() if isinstance(reads_kmer_clouds, CloudDict) else None
    if state is not None:
        idx = getattr(kmers, "_cfk_index", None)
        same_universe = idx is not None and idx[1] is state.index
        if not same_universe:
            want = np.sort(kmers_to_ints([km for km in kmers], state.k)) if len(kmers) else np.empty(0, np.uint64)
            same_universe = np.array_equal(want, to_host_u64(state.index.sorted_keys))
        if not same_universe:
            state = None
    if state is None:
        state = state_from_sets(reads_kmer_clouds, kmers=kmers)
    n_reads = len(state.r_ids)
    if min_n < 0 or (max_n is not None and max_n < 0):
        raise ValueError("Indices for islice() must be None or an integer: 0 <= x <= sys.maxsize.")
    lo_r, hi_r, _ = slice(min_n, max_n).indices(n_reads)  # itertools.islice(items, min_n, max_n), dbkr.py:92
    hi_r = max(hi_r, lo_r)
    unit_lo, unit_hi = int(state.read_unit_ptr[lo_r]), int(state.read_unit_ptr[hi_r])
    if verbose:
        print("Indexing")
        print("Inferring distances")
    kmer_index = KmerRanks(to_host_u64(state.index.sorted_keys), state.k)
    return DistCounts(state, unit_lo, unit_hi, int(min_d), int(max_d)), kmer_index


def filter_dist_tuples(dist_cnt, min_coverage, rel_threshold=0.8):
    if not isinstance(dist_cnt, DistCounts):
        raise TypeError("filter_dist_tuples needs the DistCounts handle returned by get_kmer_dist_map "
                        "(the counting and the filter are fused on the device)")
    res = dist_cnt.run(min_coverage, rel_threshold)
    edges, presorted = _sort_edges_on_device(res.edges, dist_cnt.state.index.n, dist_cnt.max_d)
    e = to_host_u32(edges, pinned=True).reshape(-1, 4)
    dist_cnt.n_increments = res.n_increments
    selected_kmers = set(to_host_u32(res.selected).tolist())
    selected_edges = EdgeList.from_rows(e, presorted=presorted)
    return selected_kmers, selected_edges


def _sort_edges_on_device(edges, n_kmers, max_d):
    """Edges (a, b, d, cnt) -> the same rows ordered by (d, a, b), sorted on the device when the three fields fit one
    63-bit key (they do unless there are more than 2^24 k-mer ids); (edges, False) otherwise (EdgeList sorts on the host)."""
    n = int(edges.shape[0])
    id_bits = max(1, int(max(n_kmers - 1, 1)).bit_length())
    d_bits = max(1, int(max(max_d, 1)).bit_length())
    if n == 0 or 2 * id_bits + d_bits > 62:
        return edges, n == 0
    import torch
    cols = edges.to(torch.int64) & 0xFFFFFFFF  # uint32 bit patterns
    key = (cols[:, 2] << (2 * id_bits)) | (cols[:, 0] << id_bits) | cols[:, 1]
    return edges[torch.argsort(key)], True


def output_results(kmer_index, min_coverage, unique_kmers_ind, dist_edges, outdir):
    if isinstance(kmer_index, KmerRanks):
        kmer_of = kmer_index.kmer_of
    else:
        rev = {index: kmer for kmer, index in kmer_index.items()}

        def kmer_of(ids):
            return [rev[int(i)] for i in ids]
    kmers_out_fn = os.path.join(outdir, f'unique_kmers_min_edge_cov_{min_coverage}.txt')
    with open(kmers_out_fn, 'w') as f:
        kmers = sorted(kmer_of(sorted(unique_kmers_ind)))
        f.write(''.join(kmer + '\n' for kmer in kmers))
    edges_out_fn = os.path.join(outdir, f'unique_edges_min_edge_cov_{min_coverage}.txt')
    if isinstance(dist_edges, EdgeList) and isinstance(kmer_index, KmerRanks):
        write_edges_native(edges_out_fn, kmer_index, dist_edges)
        return
    with open(edges_out_fn, 'w') as f:
        if isinstance(dist_edges, EdgeList):
            step = 1 << 18
            for s in range(0, len(dist_edges), step):
                sl = slice(s, s + step)
                ki, kj = kmer_of(dist_edges.i[sl]), kmer_of(dist_edges.j[sl])
                f.write(''.join(f'{d} {a} {b} {c}\n' for d, a, b, c in
                                zip(dist_edges.dist[sl].tolist(), ki, kj, dist_edges.freq[sl].tolist())))
        else:
            for t in dist_edges:
                a, b = kmer_of([t[1], t[2]])
                f.write(f'{t[0]} {a} {b} {t[3]}\n')


def write_edges_native(path, kmer_index, dist_edges, threads=0):
    """The edge file of output_results (dbkr.py:165-171) through libcfk.so's multi-threaded writer (cfk_write_edges);
    byte-identical to the f-string loop above (tests/test_result_writer.py)."""
    from . import _lib
    lib = _lib.load()
    keys = np.ascontiguousarray(kmer_index.keys_u64, dtype=np.uint64)
    if dist_edges.rows is not None:  # the device's rows, untouched
        rc = lib.cfk_write_edges_rows(os.fsencode(path), keys.ctypes.data, int(keys.size), int(kmer_index.k),
                                      dist_edges.rows.ctypes.data, len(dist_edges), int(threads))
    else:
        rc = lib.cfk_write_edges(os.fsencode(path), keys.ctypes.data, int(keys.size), int(kmer_index.k),
                                 dist_edges.dist.ctypes.data, dist_edges.i.ctypes.data, dist_edges.j.ctypes.data,
                                 dist_edges.freq.ctypes.data, len(dist_edges), int(threads))
    if rc != 0:
        raise OSError(lib.cfk_writer_last_error().decode(errors="replace"))


LAST_TIMINGS = {}  # wall-clock seconds of the stages of the last main() call (bench.py reports them as e2e_cli)


def main(argv=None):
    import time
    stamp = [("start", time.perf_counter())]
    params = parse_args(argv)
    smart_makedirs(params.outdir)

    reads_ncrf_report = LazyNCRF_Report(params.ncrf)  # Python records only if somebody asks for them
    stamp.append(("parse_report", time.perf_counter()))
    rare_kmers = get_rare_kmers(reads_ncrf_report,
                                k=params.k,
                                bottom=params.bottom,
                                top=params.top,
                                coverage=params.coverage,
                                kmer_survival_rate=params.kmer_survival_rate,
                                max_nonuniq=params.max_nonuniq,
                                verbose=params.verbose,
                                _strings=False)  # nothing below reads the strings: they are made only if asked for

    stamp.append(("ingest_count_rare", time.perf_counter()))

    reads_kmer_clouds = get_reads_kmer_clouds(reads_ncrf_report, n=1, k=params.k, genomic_kmers=rare_kmers)
    stamp.append(("clouds", time.perf_counter()))

    dist_cnt, kmer_index = get_kmer_dist_map(reads_kmer_clouds, rare_kmers,
                                             min_n=params.min_nreads, max_n=params.max_nreads,
                                             min_d=params.min_distance, max_d=params.max_distance,
                                             verbose=params.verbose)

    unique_kmers_ind, dist_edges = filter_dist_tuples(dist_cnt, min_coverage=params.min_coverage)
    stamp.append(("dist_graph_filter_sort_d2h", time.perf_counter()))

    output_results(kmer_index=kmer_index, min_coverage=params.min_coverage,
                   unique_kmers_ind=unique_kmers_ind, dist_edges=dist_edges, outdir=params.outdir)
    stamp.append(("write_files", time.perf_counter()))
    LAST_TIMINGS.clear()
    LAST_TIMINGS.update({name: t - stamp[i][1] for i, (name, t) in enumerate(stamp[1:])})
    LAST_TIMINGS["total"] = stamp[-1][1] - stamp[0][1]


if __name__ == "__main__":
    main()
