"""Row a9: the native edge-file writer (libcfk.so, csrc/result_writer.cpp) writes the bytes of the reference's
f-string loop (scripts/distance_based_kmer_recruitment.py:165-171), restated here in Python."""
import numpy as np
import pytest

from centroflye_b200 import distance_based_kmer_recruitment as dbkr
from centroflye_b200.encode import ints_to_kmers


def _python_bytes(ranks, el):
    kmers = ints_to_kmers(ranks.keys_u64, ranks.k)
    return "".join(f"{d} {kmers[a]} {kmers[b]} {c}\n" for d, a, b, c in el).encode()


@pytest.mark.parametrize("k,n_keys,n_edges", [(1, 4, 10), (5, 300, 1), (19, 5000, 70001), (31, 2000, 200003), (19, 10, 0)])
def test_native_writer_matches_python(tmp_path, k, n_keys, n_edges):
    rng = np.random.default_rng(k * 1000 + n_edges)
    keys = np.sort(rng.choice(min(1 << (2 * k), 1 << 40), size=min(n_keys, 1 << (2 * k)), replace=False).astype(np.uint64))
    if k == 31:
        keys[-1] = (1 << 62) - 1  # TTT...T: the largest k-mer
    ranks = dbkr.KmerRanks(keys, k)
    n = keys.size
    el = dbkr.EdgeList(rng.integers(1, 151, n_edges), rng.integers(0, n, n_edges), rng.integers(0, n, n_edges),
                       rng.integers(4, 3000, n_edges))
    for threads in (1, 0):
        fn = tmp_path / f"edges_{threads}.txt"
        dbkr.write_edges_native(str(fn), ranks, el, threads=threads)
        assert fn.read_bytes() == _python_bytes(ranks, el)


@pytest.mark.parametrize("k,n_keys,n_edges", [(5, 300, 1), (19, 5000, 70001), (31, 2000, 1200003), (19, 10, 0)])
def test_native_writer_from_device_rows(tmp_path, k, n_keys, n_edges):
    """EdgeList.from_rows keeps the uint32 rows (i, j, dist, freq) of cfk_pair_join; cfk_write_edges_rows writes the same
    bytes as the column form and as the Python loop; the int64 columns appear only when read.  1.2e6 edges = several
    batches of the writer's double buffering on any core count."""
    rng = np.random.default_rng(k + n_edges)
    keys = np.sort(rng.choice(min(1 << (2 * k), 1 << 40), size=min(n_keys, 1 << (2 * k)), replace=False).astype(np.uint64))
    ranks = dbkr.KmerRanks(keys, k)
    n = keys.size
    rows = np.stack([rng.integers(0, n, n_edges), rng.integers(0, n, n_edges), rng.integers(1, 151, n_edges),
                     rng.integers(4, 3000, n_edges)], 1).astype(np.uint32)
    el = dbkr.EdgeList.from_rows(rows)  # sorts by (dist, i, j) on the host
    assert el._cols is None and len(el) == n_edges
    by_cols = dbkr.EdgeList(rows[:, 2].astype(np.int64), rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64),
                            rows[:, 3].astype(np.int64))
    for threads in (1, 3, 0):
        fn = tmp_path / f"rows_{threads}.txt"
        dbkr.write_edges_native(str(fn), ranks, el, threads=threads)
        assert el._cols is None  # the writer took the rows as they are
        fn2 = tmp_path / f"cols_{threads}.txt"
        dbkr.write_edges_native(str(fn2), ranks, by_cols, threads=threads)
        assert fn.read_bytes() == fn2.read_bytes()
    if n_edges <= 100000:
        assert (tmp_path / "rows_0.txt").read_bytes() == _python_bytes(ranks, el)
    assert np.array_equal(el.dist, by_cols.dist) and np.array_equal(el.i, by_cols.i)
    assert np.array_equal(el.j, by_cols.j) and np.array_equal(el.freq, by_cols.freq)
    if n_edges:
        assert el[0] == by_cols[0] and list(el)[-1] == by_cols[len(by_cols) - 1]
    pre = dbkr.EdgeList.from_rows(el.rows, presorted=True)
    assert pre.rows is not None and np.array_equal(pre.rows, el.rows)


def test_native_writer_rows_reject_bad_ids(tmp_path):
    ranks = dbkr.KmerRanks(np.array([1, 2, 3], dtype=np.uint64), 5)
    el = dbkr.EdgeList.from_rows(np.array([[0, 3, 1, 4]], dtype=np.uint32))
    with pytest.raises(OSError, match="outside"):
        dbkr.write_edges_native(str(tmp_path / "x.txt"), ranks, el)


def test_output_results_uses_native_writer_and_sorted_edges(tmp_path):
    rng = np.random.default_rng(3)
    keys = np.sort(rng.choice(1 << 38, size=1000, replace=False).astype(np.uint64))
    ranks = dbkr.KmerRanks(keys, 19)
    el = dbkr.EdgeList(rng.integers(1, 151, 5000), rng.integers(0, 1000, 5000), rng.integers(0, 1000, 5000),
                       rng.integers(4, 40, 5000))
    assert np.array_equal(np.lexsort((el.j, el.i, el.dist)), np.arange(5000))  # sorted by (dist, i, j)
    dbkr.output_results(ranks, 4, {1, 5, 7}, el, str(tmp_path))
    assert (tmp_path / "unique_edges_min_edge_cov_4.txt").read_bytes() == _python_bytes(ranks, el)
    assert (tmp_path / "unique_kmers_min_edge_cov_4.txt").read_text().splitlines() == sorted(ranks.kmer_of([1, 5, 7]))


def test_native_writer_rejects_bad_ids(tmp_path):
    ranks = dbkr.KmerRanks(np.array([1, 2, 3], dtype=np.uint64), 5)
    el = dbkr.EdgeList(np.array([1]), np.array([0]), np.array([3]), np.array([4]))
    with pytest.raises(OSError, match="outside"):
        dbkr.write_edges_native(str(tmp_path / "x.txt"), ranks, el)


@pytest.mark.parametrize("n_kmers,max_d", [(1000, 150), (1 << 20, 150), (3, 1), ((1 << 28), 1 << 12)])
def test_device_edge_sort_matches_host_lexsort(n_kmers, max_d):
    """_sort_edges_on_device (torch; CPU tensors here) orders (a, b, d, cnt) rows by (d, a, b) like EdgeList's lexsort,
    and declines when the three fields do not fit one key."""
    import torch
    rng = np.random.default_rng(n_kmers % 1000 + max_d)
    n = 5000
    # distinct (d, a, b) triples, ids up to n_kmers - 1 (uint32 bit patterns in an int32 tensor)
    trip = np.unique(np.stack([rng.integers(0, n_kmers, n), rng.integers(0, n_kmers, n), rng.integers(1, max_d + 1, n)], 1), axis=0)
    rng.shuffle(trip)
    e = np.concatenate([trip, rng.integers(4, 99, (trip.shape[0], 1))], 1).astype(np.uint32)
    got, presorted = dbkr._sort_edges_on_device(torch.from_numpy(e.view(np.int32)), n_kmers, max_d)
    fits = 2 * max(1, (n_kmers - 1).bit_length()) + max(1, max_d.bit_length()) <= 62
    assert presorted == fits
    got = got.numpy().view(np.uint32)
    if fits:
        order = np.lexsort((e[:, 1], e[:, 0], e[:, 2]))
        assert np.array_equal(got, e[order])
    else:
        assert np.array_equal(got, e)


def test_native_writer_reports_io_errors(tmp_path):
    """A full device (the write happens on the writer thread, one batch behind the formatting) and a missing directory
    both surface as OSError with the path in the message; nothing is silently truncated."""
    import os
    rng = np.random.default_rng(5)
    keys = np.sort(rng.choice(1 << 38, size=1000, replace=False).astype(np.uint64))
    ranks = dbkr.KmerRanks(keys, 19)
    n = 300000
    rows = np.stack([rng.integers(0, 1000, n), rng.integers(0, 1000, n), rng.integers(1, 151, n), rng.integers(4, 60, n)],
                    1).astype(np.uint32)
    el = dbkr.EdgeList.from_rows(rows)
    if os.path.exists("/dev/full"):
        with pytest.raises(OSError, match="short write to /dev/full"):
            dbkr.write_edges_native("/dev/full", ranks, el, threads=2)
    with pytest.raises(OSError, match="cannot open"):
        dbkr.write_edges_native(str(tmp_path / "no_such_dir" / "x.txt"), ranks, el)
