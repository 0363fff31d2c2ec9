"""ctypes front end of oracle/hnsw_ref.c — the reference's IndexHNSWFlat restated (BENCH/TEST ONLY)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "libhnsw_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "hnsw_ref.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libhnsw_ref.so"], check=True, capture_output=True)
    return _LIB


def _L():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(_LIB))
        lib.hnsw_build.restype = C.c_void_p
        lib.hnsw_build.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64]
        lib.hnsw_search.restype = None
        lib.hnsw_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.hnsw_free.restype = None
        lib.hnsw_free.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


class IndexHNSWFlat:
    """faiss.IndexHNSWFlat(dim, M) with hnsw.efConstruction / hnsw.efSearch, METRIC_L2
    (main/create_index.py:229-234)."""

    def __init__(self, d: int, M: int = 32, efConstruction: int = 200, efSearch: int = 64, seed: int = 12345):
        self.d, self.M, self.efConstruction, self.efSearch, self.seed = d, M, efConstruction, efSearch, seed
        self._x = None
        self._g = None

    @property
    def ntotal(self) -> int:
        return 0 if self._x is None else self._x.shape[0]

    def add(self, x: np.ndarray) -> None:
        assert self._g is None, "single add() (the bench builds once)"
        self._x = np.ascontiguousarray(x, dtype=np.float32)
        self._g = _L().hnsw_build(self._x.ctypes.data, self._x.shape[0], self.d, self.M, self.efConstruction,
                                  self.seed)

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        dist = np.empty((q.shape[0], k), np.float32)
        lab = np.empty((q.shape[0], k), np.int64)
        _L().hnsw_search(self._g, q.ctypes.data, q.shape[0], k, self.efSearch, dist.ctypes.data, lab.ctypes.data)
        return dist, lab

    def __del__(self):
        if getattr(self, "_g", None):
            _L().hnsw_free(self._g)
            self._g = None
