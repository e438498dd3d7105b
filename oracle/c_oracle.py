"""ctypes wrapper of oracle/c/cfk_oracle.c (TEST INFRASTRUCTURE ONLY — the CPU checker and the timed
CPU baseline; the product path never imports this)."""
import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "cfk_oracle.c")
LIB = os.path.join(HERE, "c", "libcfk_oracle.so")

_lib = None
_p = ctypes.c_void_p
_i64 = ctypes.c_int64


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed:\n" + res.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.cfko_docfreq.restype = _i64
        _lib.cfko_docfreq.argtypes = [_p, _p, _p, _i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_p),
                                      ctypes.POINTER(_p), ctypes.POINTER(_p)]
        _lib.cfko_clouds.restype = None
        _lib.cfko_clouds.argtypes = [_p, _p, _p, _p, _i64, ctypes.c_int, _p, _i64, ctypes.c_int, _p, _p]
        _lib.cfko_dist_edges.restype = _i64
        _lib.cfko_dist_edges.argtypes = [_p, _p, _p, _i64, _i64, _i64, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                         ctypes.c_double, _i64, _i64, _i64, ctypes.c_double, ctypes.c_int,
                                         ctypes.POINTER(_p), _p, _p]
        _lib.cfko_free.argtypes = [_p]
        _lib.cfko_max_threads.restype = ctypes.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_p)


def _take(ptr, n, dtype):
    out = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)),
                                shape=(max(n, 0) * np.dtype(dtype).itemsize,)).view(dtype).copy() if n > 0 \
        else np.empty(0, dtype=dtype)
    lib().cfko_free(ptr)
    return out


def unpacked_codes(batch):
    from centroflye_b200.encode import unpack_codes
    return np.ascontiguousarray(unpack_codes(batch.packed, batch.packed.size * 16))


def docfreq(codes, batch, k, threads=1):
    """-> (keys u64, n_reads u32, n_multi u32), sorted by key."""
    pk, pr, pm = _p(), _p(), _p()
    n = lib().cfko_docfreq(_ptr(codes), _ptr(batch.read_off), _ptr(batch.read_len), batch.n_reads, k, threads,
                           ctypes.byref(pk), ctypes.byref(pr), ctypes.byref(pm))
    keys, nr, nm = _take(pk, n, np.uint64), _take(pr, n, np.uint32), _take(pm, n, np.uint32)
    o = np.argsort(keys)
    return keys[o], nr[o], nm[o]


def clouds(codes, units, k, rare_sorted, threads=1):
    """-> (unit_ptr int64[U+1], ids u32[E])."""
    U = units.n_units
    nk = np.maximum(units.unit_len.astype(np.int64) - k + 1, 0)
    kbase = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(nk, out=kbase[1:])
    tmp = np.empty(max(int(kbase[-1]), 1), dtype=np.uint32)
    cnt = np.zeros(max(U, 1), dtype=np.int32)
    rare_sorted = np.ascontiguousarray(rare_sorted, dtype=np.uint64)
    lib().cfko_clouds(_ptr(codes), _ptr(units.unit_off), _ptr(units.unit_len), _ptr(kbase), U, k, _ptr(rare_sorted),
                      rare_sorted.size, threads, _ptr(tmp), _ptr(cnt))
    ptr = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(cnt[:U], out=ptr[1:])
    ids = np.empty(int(ptr[-1]), dtype=np.uint32)
    for u in np.flatnonzero(cnt[:U]):
        ids[ptr[u]:ptr[u + 1]] = tmp[kbase[u]:kbase[u] + cnt[u]]
    return ptr, ids


def unit_last_of(units):
    ptr = units.read_unit_ptr
    return np.repeat(ptr[1:] - 1, np.diff(ptr)).astype(np.int32)


def dist_edges(unit_ptr, ids, unit_last, n_kmers, min_d, max_d, min_cov, rel_threshold=0.8, unit_lo=0, unit_hi=None,
               threads=1, sample_step=None, time_budget_s=0.0):
    """-> dict(edges u32[n,4] (a, b, d, cnt), selected ids, n_increments, n_sources_done, n_candidates)."""
    assert max_d < 65536
    unit_hi = unit_last.size if unit_hi is None else unit_hi
    selected = np.zeros(max(n_kmers, 1), dtype=np.uint8)
    stats = np.zeros(4, dtype=np.int64)
    pe = _p()
    if n_kmers == 0 or unit_hi <= unit_lo:
        return dict(edges=np.empty((0, 4), np.uint32), selected=np.empty(0, np.int64), n_increments=0,
                    n_sources_done=0, n_candidates=0)
    step = 1 if sample_step is None else sample_step
    lib().cfko_dist_edges(_ptr(unit_ptr), _ptr(ids), _ptr(unit_last), n_kmers, unit_lo, unit_hi, min_d, max_d,
                          int(min_cov), float(rel_threshold), 0, step, n_kmers, float(time_budget_s), threads,
                          ctypes.byref(pe), _ptr(selected), _ptr(stats))
    edges = _take(pe, int(stats[0]) * 4, np.uint32).reshape(-1, 4)
    return dict(edges=edges, selected=np.flatnonzero(selected[:n_kmers]), n_increments=int(stats[1]),
                n_sources_done=int(stats[2]), n_candidates=int(stats[3]))


def recruit(batch, units, k, lo, hi, max_nonuniq, min_d, max_d, min_cov, rel_threshold=0.8, threads=1,
            unit_lo=0, unit_hi=None):
    """Whole path on flat arrays; everything the GPU path returns, from the CPU."""
    codes = unpacked_codes(batch)
    keys, nr, nm = docfreq(codes, batch, k, threads)
    keep = nm <= max_nonuniq
    rare = keys[keep & (nr >= lo) & (nr <= hi)]
    unit_ptr, ids = clouds(codes, units, k, rare, threads)
    d = dist_edges(unit_ptr, ids, unit_last_of(units), rare.size, min_d, max_d, min_cov, rel_threshold,
                   unit_lo=unit_lo, unit_hi=unit_hi, threads=threads)
    return dict(all_keys=keys[keep], all_counts=nr[keep], rare=rare, unit_ptr=unit_ptr, ids=ids, **d)


def coprime_step(n):
    step = max(1, int(n * 0.6180339887))
    while np.gcd(step, max(n, 1)) != 1:
        step += 1
    return step


def timed_sample(batch, units, params, band, bounded_s=15.0, threads=1):
    """CPU baseline for bench.py: stages A and B on ALL reads, stage C/D on a pseudo-random sample of source
    k-mers bounded by ``bounded_s`` seconds and extrapolated by the share of sources visited."""
    k = params["k"]
    lo, hi = band
    t0 = time.perf_counter()
    codes = unpacked_codes(batch)
    t1 = time.perf_counter()
    keys, nr, nm = docfreq(codes, batch, k, threads)
    rare = keys[(nm <= params["max_nonuniq"]) & (nr >= lo) & (nr <= hi)]
    t2 = time.perf_counter()
    unit_ptr, ids = clouds(codes, units, k, rare, threads)
    t3 = time.perf_counter()
    n = int(rare.size)
    d = dist_edges(unit_ptr, ids, unit_last_of(units), n, params["min_d"], params["max_d"], params["min_coverage"],
                   threads=threads, sample_step=coprime_step(n), time_budget_s=bounded_s)
    t4 = time.perf_counter()
    frac = d["n_sources_done"] / n if n else 1.0
    t_cd = (t4 - t3) / frac if frac > 0 else float("inf")
    total = (t2 - t1) + (t3 - t2) + t_cd
    return {"value": batch.n_bases / total, "unit": "read-bases/s", "cores": threads, "kind": "port",
            "ms": total * 1e3,
            "sample": f"oracle/c (C restatement of the reference's Python): stage A {t2 - t1:.2f}s and stage B "
                      f"{t3 - t2:.2f}s on all {batch.n_bases} read bases; stage C/D on {d['n_sources_done']} of {n} "
                      f"source k-mers in {t4 - t3:.2f}s ({d['n_increments']} pair increments), extrapolated to all "
                      f"sources ({t_cd:.1f}s); host has {os.cpu_count()} cpus",
            "stage_s": {"A": t2 - t1, "B": t3 - t2, "CD_sample": t4 - t3, "CD_extrapolated": t_cd},
            "sources_done": d["n_sources_done"], "sources": n}
