"""The k-mer analysis of the reference's ``scripts/unit_extractor.py`` with the grouping of k-mer
positions done on the device: ``get_repetitive_kmers`` (:23-31), ``get_convolution`` (:33-41), ``get_period_info``
(:43-80), ``get_hook_kmer`` (:83-91), ``split_by_hook`` (:94-105) -- same names, arguments and return values.  The rest of
that script (Flye polishing of the splits, the histogram plot) stays with the reference.

One sort key (k-mer << position bits | position) per k-mer start is sorted on the device (cfk_kmer_position_keys,
cfk_sort_u64, cfk_adjacent_gaps): equal k-mers end up neighbours in position order, which is both the position list of
``get_repetitive_kmers`` and, by differences of neighbours, the gap list of ``get_convolution``.  The dictionaries the
callers expect are built from those arrays; ``get_period_info`` is a sequential sweep over the sorted gaps and runs on
the host as written in the reference.
"""
import numpy as np

from . import _lib
from .encode import ascii_to_codes, check_k, ints_to_kmers, pack_codes


class RepKmers(dict):
    """kmer -> increasing positions (only k-mers occurring more than once), in order of first occurrence like the
    reference's dict; remembers the device-sorted arrays so that get_convolution need not regroup strings."""
    _cfk = None


def _sorted_occurrences(seq, k):
    from .engine import default_engine
    k = check_k(k)
    eng = default_engine()
    t = eng.torch
    codes = ascii_to_codes(seq)
    n = int(codes.size) - k + 1
    if n <= 0:
        z = np.zeros(0, dtype=np.uint64)
        return z, z.astype(np.int64), np.zeros(0, dtype=np.uint32)
    pos_bits = max(1, int(codes.size).bit_length())
    if 2 * k + pos_bits > 63:
        raise ValueError(f"sequence of {codes.size} bases is too long for k = {k} (2k + position bits must fit 63)")
    packed = pack_codes(np.concatenate([codes, np.zeros(64, dtype=np.uint8)]))  # + the extraction window's overhang
    d_packed = eng._to_dev(packed.view(np.int32))
    keys = eng._empty(n, t.int64)
    _lib.call("cfk_kmer_position_keys", eng._p(d_packed), int(codes.size), k, pos_bits, eng._p(keys), eng._stream())
    _lib.call("cfk_sort_u64", eng._p(keys), n, eng._stream())
    gaps = eng._empty(n, t.int32)
    _lib.call("cfk_adjacent_gaps", eng._p(keys), n, pos_bits, eng._p(gaps), eng._stream())
    hk = keys[:n].cpu().numpy().view(np.uint64)
    return hk >> np.uint64(pos_bits), (hk & np.uint64((1 << pos_bits) - 1)).astype(np.int64), gaps[:n].cpu().numpy().view(np.uint32)


def get_repetitive_kmers(seq, k):
    kmer, pos, gaps = _sorted_occurrences(seq, k)
    out = RepKmers()
    if kmer.size == 0:
        return out
    first = np.flatnonzero(np.concatenate([[True], kmer[1:] != kmer[:-1]]))  # start of every k-mer's run
    size = np.diff(np.concatenate([first, [kmer.size]]))
    keep = size > 1
    first, size = first[keep], size[keep]
    order = np.argsort(pos[first], kind="stable")  # the reference's dict is in order of first occurrence
    names = ints_to_kmers(kmer[first[order]], k)
    for name, f, s in zip(names, first[order].tolist(), size[order].tolist()):
        out[name] = pos[f:f + s].tolist()
    out._cfk = dict(k=k, first=first[order], size=size[order], pos=pos, gaps=gaps, names=names)
    return out


def get_convolution(rep_kmers):
    conv, union_conv = {}, []
    c = getattr(rep_kmers, "_cfk", None)
    if c is not None and len(c["names"]) == len(rep_kmers):
        for name, f, s in zip(c["names"], c["first"].tolist(), c["size"].tolist()):
            conv[name] = np.sort(c["gaps"][f + 1:f + s]).tolist()  # gaps[f] = 0 marks the run's first occurrence
        in_run = np.zeros(c["gaps"].size, dtype=bool)
        for f, s in zip(c["first"].tolist(), c["size"].tolist()):
            in_run[f + 1:f + s] = True
        union_conv = np.sort(c["gaps"][in_run]).tolist()
        return conv, union_conv
    for kmer in rep_kmers:  # a plain dict from elsewhere: the reference's own arithmetic
        pos = rep_kmers[kmer]
        conv[kmer] = sorted(y - x for x, y in zip(pos[:-1], pos[1:]))
        union_conv += conv[kmer]
    union_conv.sort()
    return conv, union_conv


def get_period_info(conv, bin_size):
    """unit_extractor.py:43-80: a two-pointer sweep over the SORTED gap list; windows of width 2 * bin_size, their
    median as the period, the largest window per period kept."""
    if len(conv) == 0:
        return [], [], None, None
    periods2bin_convs, bin_convs2periods = {}, {}
    left, right = 0, 0
    best_l, best_r = 0, 0
    n = len(conv)
    while right < n:
        while right < n and conv[right] - conv[left] <= 2 * bin_size:
            right += 1
        mid = left + (right - left) // 2
        period = (conv[mid] + conv[mid - 1]) // 2 if (right - left) % 2 == 0 else conv[mid]
        width = right - left
        if period not in periods2bin_convs or width > periods2bin_convs[period]:
            bin_convs2periods[width] = period
            if period in periods2bin_convs and width > periods2bin_convs[period]:
                bin_convs2periods.pop(periods2bin_convs[period], None)
            periods2bin_convs[period] = width
        if width > best_r - best_l:
            best_l, best_r = left, right
        left += 1
    bin_convs, periods = zip(*sorted(bin_convs2periods.items(), reverse=True))
    return periods, bin_convs, conv[best_l], conv[best_r - 1]


def get_hook_kmer(conv, bin_left, bin_right):
    """The k-mer with the most gaps inside [bin_left, bin_right]; the first such k-mer in dict order on ties (:83-91)."""
    from bisect import bisect_left, bisect_right
    hook_kmer, max_tandem_index = None, 0
    for kmer, dist in conv.items():
        tandem_index = bisect_right(dist, bin_right) - bisect_left(dist, bin_left)
        if tandem_index > max_tandem_index:
            hook_kmer, max_tandem_index = kmer, tandem_index
    return hook_kmer


def split_by_hook(seq, hook):
    """Pieces of seq between consecutive occurrences of the hook, keyed split_<start>_<end> (:94-105)."""
    hook_pos, start = [], seq.find(hook)
    while start >= 0:
        hook_pos.append(start)
        start = seq.find(hook, start + 1)
    return {f'split_{a}_{b}': seq[a:b] for a, b in zip(hook_pos[:-1], hook_pos[1:])}
