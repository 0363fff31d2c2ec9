"""ctypes/numpy front end of oracle/b2k_oracle.c (TEST INFRASTRUCTURE ONLY — see package doc)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libb2k_oracle.so"
_lib = None


def build_oracle(force: bool = False) -> Path:
    """Compile oracle/b2k_oracle.c with the committed Makefile (gcc, seconds)."""
    src = _HERE / "b2k_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libb2k_oracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


def _L():
    global _lib
    if _lib is None:
        build_oracle()
        lib = C.CDLL(str(_LIB_PATH))
        lib.orc_sumsq32.restype = C.c_float
        lib.orc_sumsq32.argtypes = [C.c_void_p, C.c_int]
        lib.orc_dot_exact.restype = C.c_float
        lib.orc_bf16_rne.restype = C.c_uint16
        lib.orc_bf16_rne.argtypes = [C.c_float]
        lib.orc_pack.restype = None
        lib.orc_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_normalize_l2.restype = None
        lib.orc_normalize_l2.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        lib.orc_search_exact.restype = None
        lib.orc_search_exact.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
        lib.orc_scores_f64.restype = None
        lib.orc_scores_f64.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        lib.orc_merge_topk.restype = None
        lib.orc_merge_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_synth_rows.restype = None
        lib.orc_synth_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_uint64, C.c_int32, C.c_float, C.c_uint32,
                                       C.c_int32, C.c_uint64, C.c_float]
        lib.orc_synth_query_source.restype = C.c_int64
        lib.orc_synth_query_source.argtypes = [C.c_uint64, C.c_int64, C.c_int64]
        lib.orc_num_threads.restype = C.c_int32
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(_L().orc_num_threads())


def dims_total(dims) -> tuple[int, int]:
    d = int(sum(dims))
    return d, (d + 63) // 64 * 64


def bf16_rne(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns (uint16), round-to-nearest-even."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = 0x7FFF
    return r


def bf16_to_f32(h: np.ndarray) -> np.ndarray:
    return (h.astype(np.uint32) << 16).view(np.float32)


def sumsq32(x: np.ndarray) -> np.float32:
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    return np.float32(_L().orc_sumsq32(_p(x), x.size))


def normalize_l2(x: np.ndarray) -> np.ndarray:
    """faiss.normalize_L2 restated (search_from_image.py:322); returns a normalised copy."""
    y = np.array(x, dtype=np.float32, order="C", copy=True)
    assert y.ndim == 2
    _L().orc_normalize_l2(_p(y), y.shape[0], y.shape[1])
    return y


def pack(tables: list[np.ndarray], normalize: bool = True):
    """Spec P.  tables[t]: [n, d_t] fp32.  Returns dict(f32, bf16, norm2, stats)."""
    tabs = [np.ascontiguousarray(t, dtype=np.float32) for t in tables]
    n = tabs[0].shape[0]
    dims = np.array([t.shape[1] for t in tabs], dtype=np.int32)
    D, Dp = dims_total(dims)
    ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    f32 = np.empty((n, D), np.float32)
    b16 = np.empty((n, Dp), np.uint16)
    n2 = np.empty((n,), np.float32)
    stats = np.zeros((2,), np.float32)
    _L().orc_pack(C.cast(ptrs, C.c_void_p), _p(dims), len(tabs), n, int(normalize), _p(f32),
                  _p(b16), _p(n2), _p(stats))
    return {"f32": f32, "bf16": b16, "norm2": n2, "stats": stats}


def search_exact(db_f32: np.ndarray, q: np.ndarray, k: int, norm2: np.ndarray | None = None,
                 base_offset: int = 0):
    """Spec R exact top-k.  Returns (dist [nq,k] f32, labels [nq,k] i64, ip [nq,k] f32)."""
    db = np.ascontiguousarray(db_f32, dtype=np.float32)
    qq = np.ascontiguousarray(q, dtype=np.float32)
    n, D = db.shape
    nq = qq.shape[0]
    assert qq.shape[1] == D
    ip = np.empty((nq, k), np.float32)
    dist = np.empty((nq, k), np.float32)
    lab = np.empty((nq, k), np.int64)
    n2 = None if norm2 is None else np.ascontiguousarray(norm2, dtype=np.float32)
    _L().orc_search_exact(_p(db), _p(n2), n, D, _p(qq), nq, k, base_offset, _p(ip), _p(dist),
                          _p(lab))
    return dist, lab, ip


def scores_f64(db_f32: np.ndarray, q1: np.ndarray) -> np.ndarray:
    db = np.ascontiguousarray(db_f32, dtype=np.float32)
    qq = np.ascontiguousarray(q1, dtype=np.float32).ravel()
    out = np.empty((db.shape[0],), np.float64)
    _L().orc_scores_f64(_p(db), db.shape[0], db.shape[1], _p(qq), _p(out))
    return out


def merge_topk(ip: np.ndarray, dist: np.ndarray, labels: np.ndarray):
    """[n_lists, nq, k] x3 -> (dist, labels, ip) [nq, k]; higher ip first, then lower offset."""
    ip = np.ascontiguousarray(ip, np.float32)
    dist = np.ascontiguousarray(dist, np.float32)
    labels = np.ascontiguousarray(labels, np.int64)
    nl, nq, k = ip.shape
    o_ip = np.empty((nq, k), np.float32)
    o_d = np.empty((nq, k), np.float32)
    o_l = np.empty((nq, k), np.int64)
    _L().orc_merge_topk(_p(ip), _p(dist), _p(labels), nl, nq, k, _p(o_ip), _p(o_d), _p(o_l))
    return o_d, o_l, o_ip


def _synth(dims, n, first, total_rows, seed, n_clusters, sigma, abs_mask, query_mode, qseed,
           sigma_q):
    dims = np.asarray(dims, dtype=np.int32)
    tabs = [np.empty((n, int(d)), np.float32) for d in dims]
    ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    _L().orc_synth_rows(C.cast(ptrs, C.c_void_p), _p(dims), len(tabs), n, first, total_rows,
                        seed, n_clusters, sigma, abs_mask, query_mode, qseed, sigma_q)
    return tabs


def synth_rows(dims, n, first=0, total_rows=None, seed=0xC0FFEE, n_clusters=4096, sigma=0.3,
               abs_mask=1):
    """Spec G raw per-table DB rows [first, first+n) (un-normalised)."""
    total_rows = n + first if total_rows is None else total_rows
    return _synth(dims, n, first, total_rows, seed, n_clusters, sigma, abs_mask, 0, 0, 0.0)


def synth_queries(dims, nq, total_rows, seed=0xC0FFEE, n_clusters=4096, sigma=0.3, abs_mask=1,
                  qseed=0x5EED, sigma_q=0.05, first=0):
    """Spec G queries: noisy copies of seeded DB rows; every part is unit-normalised like an
    extractor output (Spec P), the parts are concatenated and the whole vector normalised
    (search_from_image.py:305-322)."""
    tabs = _synth(dims, nq, first, total_rows, seed, n_clusters, sigma, abs_mask, 1, qseed,
                  sigma_q)
    return normalize_l2(pack(tabs, normalize=True)["f32"])


def synth_query_source(qseed: int, query_index: int, total_rows: int) -> int:
    return int(_L().orc_synth_query_source(qseed, query_index, total_rows))
