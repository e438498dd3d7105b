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
