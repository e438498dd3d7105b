"""RareKmerSet (centroflye_b200/distance_based_kmer_recruitment.py): the set[str] get_rare_kmers returns.  It carries
the device copy of the rare set, drops it on any in-place change, and -- inside main() only -- may hold its strings
pending until somebody looks at them.  Host logic only, no GPU."""
import pickle

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401

from centroflye_b200.distance_based_kmer_recruitment import RareKmerSet
from centroflye_b200.encode import kmers_to_ints

KMERS = ["ACGTA", "CCCCC", "TTTTT"]


def pending():
    r = RareKmerSet()
    r._cfk_k = 5
    r._cfk_pending = np.sort(kmers_to_ints(KMERS, 5))
    r._cfk_index = ("engine", "index")
    return r


def test_reads_fill_the_strings_and_keep_the_device_copy():
    looks = [lambda r: len(r) == 3, lambda r: "CCCCC" in r and "AAAAA" not in r, lambda r: sorted(r) == sorted(KMERS),
             lambda r: r == set(KMERS), lambda r: (r | {"GGGGG"}) == set(KMERS) | {"GGGGG"},
             lambda r: ({"GGGGG"} | r) == set(KMERS) | {"GGGGG"}, lambda r: r.copy() == set(KMERS),
             lambda r: r.issuperset({"ACGTA"}) and r.isdisjoint({"GGGGG"}), lambda r: (r - {"ACGTA"}) == {"CCCCC", "TTTTT"},
             lambda r: "CCCCC" in repr(r), lambda r: pickle.loads(pickle.dumps(r)) == set(KMERS),
             lambda r: r.materialize() is r and set(r) == set(KMERS)]
    for look in looks:
        r = pending()
        assert look(r)
        assert r._cfk_pending is None and r._cfk_index == ("engine", "index")
        assert set.__len__(r) == 3


@pytest.mark.parametrize("change", [lambda r: r.add("GGGGG"), lambda r: r.discard("CCCCC"), lambda r: r.remove("CCCCC"),
                                    lambda r: r.pop(), lambda r: r.clear(), lambda r: r.update({"GGGGG"}),
                                    lambda r: r.difference_update({"CCCCC"}), lambda r: r.intersection_update({"CCCCC"}),
                                    lambda r: r.symmetric_difference_update({"CCCCC"}),
                                    lambda r: r.__ior__({"GGGGG"}), lambda r: r.__isub__({"CCCCC"})])
def test_in_place_changes_fill_first_and_drop_the_device_copy(change):
    r, plain = pending(), set(KMERS)
    change(r)
    assert r._cfk_index is None and r._cfk_pending is None
    if change.__code__.co_names[-1] != "pop":  # pop() takes an arbitrary element
        change(plain)
        assert set.__eq__(r, plain)
    else:
        assert set.__len__(r) == 2


def test_a_filled_set_behaves_like_a_set():
    r = RareKmerSet(KMERS)
    assert r._cfk_pending is None and len(r) == 3 and r == set(KMERS) and isinstance(r, set)
    with pytest.raises(TypeError):
        hash(r)
