import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_ready():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without CUDA skips the gpu-marked tests instead of failing them; on a GPU box nothing
    is skipped (a missing libcfk.so fails loudly there), and the product code itself still raises without CUDA."""
    if _gpu_ready():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases():
    return sorted(d for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "report.ncrf.gz")))


def golden_points(case):
    d = os.path.join(GOLDEN, case)
    return sorted(int(f[1:-5]) for f in os.listdir(d) if f.startswith("p") and f.endswith(".json"))


def all_case_points():
    return [(c, p) for c in golden_cases() for p in golden_points(c)]


class GoldenCase:
    """One tests/golden/<case> directory: the NCRF report + what the reference computed from it."""

    def __init__(self, name, tmpdir):
        self.name = name
        self.dir = os.path.join(GOLDEN, name)
        self.report_path = os.path.join(str(tmpdir), f"{name}.ncrf")
        with gzip.open(os.path.join(self.dir, "report.ncrf.gz"), "rb") as f, open(self.report_path, "wb") as g:
            g.write(f.read())
        with gzip.open(os.path.join(self.dir, "parsed.json.gz"), "rb") as f:
            self.parsed = json.loads(f.read())

    def point(self, i):
        with open(os.path.join(self.dir, f"p{i}.json")) as f:
            meta = json.load(f)
        arrays = dict(np.load(os.path.join(self.dir, f"p{i}.npz")))
        return meta, arrays


_cache = {}


@pytest.fixture(scope="session")
def golden(tmp_path_factory):
    def get(name):
        if name not in _cache:
            _cache[name] = GoldenCase(name, tmp_path_factory.mktemp("golden"))
        return _cache[name]
    return get


def clouds_to_csr(clouds, order, k):
    """dict r_id -> list of iterables of k-mer strings -> (units_per_read, unit_ptr, sorted u64 per unit)."""
    from centroflye_b200.encode import kmers_to_ints
    n_units = np.array([len(clouds[r]) for r in order], dtype=np.int64)
    sizes = np.array([len(u) for r in order for u in clouds[r]], dtype=np.int64)
    ptr = np.zeros(sizes.size + 1, dtype=np.int64)
    np.cumsum(sizes, out=ptr[1:])
    vals = kmers_to_ints([kmer for r in order for u in clouds[r] for kmer in u], k)
    for lo, hi in zip(ptr[:-1], ptr[1:]):
        vals[lo:hi].sort()
    return n_units, ptr, vals
