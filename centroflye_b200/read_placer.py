"""Drop-in for the reference's ``scripts/read_placer.py``: same ``ReadPlacer`` class, same command
line, same ``read_positions.csv`` -- with the cloud contig, ``kmers2pos`` and the mapping scores kept on the device
(csrc/placer.cu) instead of nested Python dicts of strings.

What stays on the host is the greedy loop of ``ReadPlacer.add_reads`` (read_placer.py:58-94): one read is placed per
iteration, and which one depends on the scores after the previous placement.  Per iteration the device (1) adds the
newly frequent (k-mer, position) pairs to the scores of every unplaced read holding the k-mer (update_mapping_scores,
cloud_contig.py:87-95), (2) picks the best (read, offset) under the rule of read_placer.py:61-79, (3) adds the winner's
clouds to the contig (CloudContig.add_read, cloud_contig.py:26-41).  The clouds never become Python strings.

The reference's quirks are kept: the list an ``add_reads`` call starts from holds every position of every k-mer that is
frequent at SOME position (read_placer.py:54-57), later iterations only the pairs that just became frequent; scores of
reads placed earlier keep growing and are ignored; unplaced reads end the file as ``r_id None`` (in the reference in set
order, here in record order).
"""
import argparse
import os
from collections import defaultdict

import numpy as np

from . import _lib
from ._lib import CfkError
from .ncrf_parser import NCRF_Report
from .read_kmer_cloud import CloudDict, filter_reads_kmer_clouds, get_reads_kmer_clouds, state_from_sets


def smart_makedirs(dirname):
    os.makedirs(dirname, exist_ok=True)


class CloudContig:
    """The host face of the device contig: what read_placer.py reads of cloud_contig.CloudContig (max_pos,
    read_positions, coverage, min_cloud_kmer_freq); the counters themselves live in a device table."""

    def __init__(self, min_cloud_kmer_freq):
        self.max_pos = 0
        self.min_cloud_kmer_freq = max(1, min_cloud_kmer_freq)
        self.read_positions = {}
        self.coverage = defaultdict(int)
        self._dev = None  # bound to a CloudState by the first add_read

    def _bind(self, state):
        if self._dev is not None:
            if self._dev["state_id"] != id(state.index):
                raise CfkError("CloudContig: all reads of one contig must come from one get_reads_kmer_clouds() result")
            return self._dev
        eng, t = state.engine, state.engine.torch
        cap = 2 * int(state.csr.n_entries) + 1024
        self._dev = dict(state_id=id(state.index), cap=cap,
                         keys=t.full((cap,), -1, dtype=t.int64, device=eng.device), cnt=eng._zeros(cap, t.int32),
                         freq=eng._zeros(max(state.index.n, 1), t.uint8))
        return self._dev

    def add_read(self, state, read_index, r_id, position, pairs=None, max_pairs=0, counters=None):
        """CloudContig.add_read (cloud_contig.py:26-41) for read `read_index` of `state`; the newly frequent pairs go
        to `pairs` on the device (counters[0] of them) when given."""
        eng = state.engine
        dev = self._bind(state)
        u0, u1 = int(state.read_unit_ptr[read_index]), int(state.read_unit_ptr[read_index + 1])
        self.read_positions[r_id] = position
        for i in range(u1 - u0):
            self.coverage[i + position] += 1
        if u1 > u0:
            self._max_key = max(getattr(self, "_max_key", -1), position + (u1 - u0) - 1)
        self.max_pos = max(getattr(self, "_max_key", -1), 0)  # update_max_pos: max(clouds.keys()) or 0
        if u1 == u0:
            return
        counters = eng._counters() if counters is None else counters
        sizes = state.host_unit_sizes()
        n_entries = int(sizes[u0:u1].sum())
        _lib.call("cfk_placer_add_read", eng._p(state.csr.unit_ptr), eng._p(state.csr.ids), u0, u1 - u0, n_entries,
                  int(position), int(self.min_cloud_kmer_freq), eng._p(dev["keys"]), eng._p(dev["cnt"]), dev["cap"],
                  eng._p(dev["freq"]), eng._p(pairs), int(max_pairs), eng._p(counters), eng._stream())


class ReadPlacer:
    def __init__(self, params):
        self.params = params
        self.ncrf_report = NCRF_Report(params.ncrf)
        self.cloud_contig = CloudContig(params.min_cloud_kmer_freq)
        if params.genomic_kmers is not None:
            kmers = []
            with open(params.genomic_kmers) as f:
                for line in f:
                    kmers.append(line.strip())
            self.genomic_kmers = set(kmers)
        else:
            self.genomic_kmers = None
        smart_makedirs(params.outdir)
        self.position_outfile = os.path.join(self.params.outdir, 'read_positions.csv')

    def reset_cloud_contig(self):
        self.cloud_contig = CloudContig(self.params.min_cloud_kmer_freq)

    @staticmethod
    def _state(reads_kmer_clouds):
        state = reads_kmer_clouds.device_state() if isinstance(reads_kmer_clouds, CloudDict) else None
        if state is None:  # a foreign dict, or sets somebody edited on the host: convert once
            state = getattr(reads_kmer_clouds, "_cfk_placer_state", None) or state_from_sets(reads_kmer_clouds)
            try:
                reads_kmer_clouds._cfk_placer_state = state
            except AttributeError:
                pass
        return state

    def add_prefix_reads(self, prefix_reads, reads_kmer_clouds):
        state = self._state(reads_kmer_clouds)
        index = {r_id: i for i, r_id in enumerate(state.r_ids)}
        with open(self.position_outfile, 'w') as f:
            for r_id in prefix_reads:
                self.cloud_contig.add_read(state, index[r_id], r_id, position=0)
                print(r_id, 0, file=f)

    def add_reads(self, reads, reads_kmer_clouds, min_unit, min_inters, min_prop=3):
        state = self._state(reads_kmer_clouds)
        grow = 1
        while True:
            snapshot = None
            dev = self.cloud_contig._dev
            if dev is not None:
                snapshot = (dev["keys"].clone(), dev["cnt"].clone(), dev["freq"].clone(), dict(self.cloud_contig.read_positions),
                            dict(self.cloud_contig.coverage), getattr(self.cloud_contig, "_max_key", -1))
            lines = self._add_reads_once(list(reads), state, min_unit, min_inters, min_prop, grow)
            if lines is not None:
                break
            grow *= 4  # a score table was too small: put the contig back and go again with more room
            if grow > 4 ** 6:
                raise CfkError("read placer: score tables overflowed repeatedly")
            if snapshot is not None:
                dev["keys"].copy_(snapshot[0]); dev["cnt"].copy_(snapshot[1]); dev["freq"].copy_(snapshot[2])
                self.cloud_contig.read_positions = snapshot[3]
                self.cloud_contig.coverage = defaultdict(int, snapshot[4])
                self.cloud_contig._max_key = snapshot[5]
                self.cloud_contig.max_pos = max(snapshot[5], 0)
            else:
                self.reset_cloud_contig()
        with open(self.position_outfile, 'a') as f:
            f.write("".join(lines))

    def _add_reads_once(self, reads, state, min_unit, min_inters, min_prop, grow):
        eng, t = state.engine, state.engine.torch
        cc = self.cloud_contig
        dev = cc._bind(state)
        R, U = len(state.r_ids), int(state.csr.n_units)
        index = {r_id: i for i, r_id in enumerate(state.r_ids)}
        per_read = np.diff(state.read_unit_ptr)
        if R >= (1 << 24) or (per_read.size and int(per_read.max()) >= (1 << 16)):
            raise CfkError("read placer: more than 2^24 reads or 2^16 units in one read")
        sel = np.zeros(R, dtype=np.uint8)
        for r_id in reads:
            sel[index[r_id]] = 1
        order = sorted(range(R), key=lambda i: state.r_ids[i])  # `r_id < best_read` is a string comparison
        rank = np.empty(R, dtype=np.int32)
        rank[order] = np.arange(R, dtype=np.int32)
        d_sel, d_unused, d_rank = eng._to_dev(sel), eng._to_dev(sel.copy()), eng._to_dev(rank)
        d_unit_read = eng._to_dev(np.repeat(np.arange(R, dtype=np.int32), per_read))
        d_first = eng._to_dev(state.read_unit_ptr[:-1].astype(np.int64))
        occ_ptr, occ, _ = eng.build_occurrences(state.csr, state.index.n)
        n_sel_entries = int(state.host_unit_sizes()[np.repeat(sel.astype(bool), per_read)].sum()) if U else 0
        cap1 = grow * max(1 << 16, 16 * n_sel_entries)
        cap2 = grow * max(1 << 14, 4 * n_sel_entries)
        max_pairs = max(1 << 16, 2 * dev["cap"])
        m1 = t.full((cap1,), -1, dtype=t.int64, device=eng.device)
        m2k = t.full((cap2,), -1, dtype=t.int64, device=eng.device)
        m2v = eng._zeros(cap2, t.int64)
        pairs = eng._empty(2 * max_pairs, t.int32)
        counters = eng._counters()
        n_blocks = int(eng.lib.cfk_placer_best_blocks())
        best_out = eng._empty(3 * n_blocks, t.int64)
        _lib.call("cfk_placer_initial_pairs", eng._p(dev["keys"]), dev["cap"], eng._p(dev["freq"]), eng._p(pairs), max_pairs,
                  eng._p(counters), eng._stream())
        unused = [r_id for r_id in reads]
        unused_set = set(unused)
        n_reads = len(unused_set)
        lines = []
        while unused_set:
            _lib.call("cfk_placer_update", eng._p(pairs), eng._p(counters), max_pairs, eng._p(occ_ptr), eng._p(occ),
                      eng._p(d_unit_read), eng._p(d_first), eng._p(d_sel), eng._p(m1), cap1, eng._p(m2k), eng._p(m2v), cap2,
                      eng._p(counters), eng._stream())
            _lib.call("cfk_placer_best", eng._p(m2k), eng._p(m2v), cap2, eng._p(d_unused), eng._p(d_rank), int(min_unit),
                      int(min_inters), int(min_prop), eng._p(best_out), eng._stream())
            host = t.cat([best_out, counters]).cpu().numpy()
            if int(host[3 * n_blocks + 1]):
                return None  # overflow: the caller repeats the call with larger tables
            cand = host[: 3 * n_blocks].view(np.uint64).reshape(n_blocks, 3)
            score, off, rk = cand[:, 0], cand[:, 1] & np.uint64(0xFFFFFFFF), cand[:, 1] >> np.uint64(32)
            live = np.flatnonzero(score > 0)
            if live.size == 0:
                print(f"Unused reads {len(unused_set)}, {n_reads}, {len(unused_set) / n_reads}")
                lines += [f"{r_id} None\n" for r_id in unused if r_id in unused_set]
                return lines
            pick = live[np.lexsort((-rk[live].astype(np.int64), off[live], score[live]))[-1]]
            best_read = state.r_ids[int(cand[pick, 2] & np.uint64(0xFFFFFFFF))]
            best_position = int(off[pick])
            best_score = (int(score[pick] >> np.uint64(32)), int(score[pick] & np.uint64(0xFFFFFFFF)))
            print(best_score, best_position, best_read)
            print("")
            lines.append(f"{best_read} {best_position} {best_score[0]} {best_score[1]}\n")
            counters.zero_()
            cc.add_read(state, index[best_read], best_read, best_position, pairs=pairs, max_pairs=max_pairs, counters=counters)
            d_unused[index[best_read]] = 0
            unused_set.remove(best_read)
        return lines

    def run(self):
        left_PT_reads, FT_reads, right_PT_reads = self.ncrf_report.classify(large_threshold=self.params.prefix_threshold)
        print(f'Left: {len(left_PT_reads)}')
        print(f'FT: {len(FT_reads)}')
        print(f'Right: {len(right_PT_reads)}')
        print("Reading kmer clouds from reads")
        reads_kmer_clouds = get_reads_kmer_clouds(self.ncrf_report, n=self.params.n_motif, k=self.params.k_cloud,
                                                  genomic_kmers=self.genomic_kmers)
        print("Filtering kmer clouds from reads")
        reads_kmer_clouds = filter_reads_kmer_clouds(reads_kmer_clouds, min_mult=self.params.min_kmer_mult)
        print("Adding prefix reads")
        self.add_prefix_reads(left_PT_reads, reads_kmer_clouds)
        print(self.cloud_contig.max_pos)
        print("Adding inner reads")
        self.add_reads(FT_reads, reads_kmer_clouds, min_unit=self.params.min_unit, min_inters=self.params.min_inters)
        print(self.cloud_contig.max_pos)
        print("\nNow adding suffix reads")
        self.add_reads(right_PT_reads, reads_kmer_clouds, min_unit=self.params.min_unit, min_inters=self.params.min_inters)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--ncrf', help='NCRF report on reads', required=True)
    parser.add_argument('--genomic-kmers', help='Unique genomic kmers if known', required=True)
    parser.add_argument('--n-motif', help='Number of motifs stuck together', default=1, type=int)
    parser.add_argument('--k-cloud', help='Size of k-mer for k-mer cloud', default=19, type=int)
    parser.add_argument('--min-cloud-kmer-freq', help='Minimal frequency of a kmer in the cloud', default=2, type=int)
    parser.add_argument('--min-kmer-mult', help='Minimal frequency of a kmer in input', default=2, type=int)
    parser.add_argument('--min-unit', help='Score[0]', default=2, type=int)
    parser.add_argument('--min-inters', help='Score[1]', default=10, type=int)
    parser.add_argument('--prefix-threshold', help='Min pre/suffix length for read classification', default=50000, type=int)
    parser.add_argument('--outdir', help='Output directory', required=True)
    params = parser.parse_args(argv)
    ReadPlacer(params).run()


if __name__ == "__main__":
    main()
