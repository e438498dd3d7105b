"""CPU-side checks of the C ABI: libcfk.so builds for sm_100a, loads without a GPU, and exports every
symbol include/cfk.h declares; the ctypes table in centroflye_b200/_lib.py covers exactly those."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "cfk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfk_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from centroflye_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(lib_path)
    for name in names:
        assert hasattr(lib, name), name
    lib.cfk_abi_version.restype = ctypes.c_int
    assert lib.cfk_abi_version() == 1
    lib.cfk_scan_scratch_elems.restype = ctypes.c_int64
    lib.cfk_scan_scratch_elems.argtypes = [ctypes.c_int64]
    assert lib.cfk_scan_scratch_elems(5000) >= 3


def test_ctypes_table_matches_header():
    from centroflye_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.load()


def test_argument_validation_without_gpu(lib_path):
    """Host-side argument checks run before any CUDA call, so they work on a CPU-only box."""
    from centroflye_b200 import _lib
    lib = _lib.load()
    with pytest.raises(_lib.CfkError, match="k must be"):
        _lib.call("cfk_docfreq_count", None, None, None, None, 1, 32, None, 10, None, 1, None)
    with pytest.raises(_lib.CfkError, match="min_d"):
        _lib.call("cfk_pair_candidates", None, None, None, None, None, None, 5, 10, 0, 10, 1, -1, 5, 1, None, 0, None, 1,
                  None)
    assert b"min_d" in lib.cfk_last_error()


def test_built_for_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from centroflye_b200._lib import CfkError
    from centroflye_b200.engine import Engine
    with pytest.raises(CfkError, match="no CPU fallback"):
        Engine()
