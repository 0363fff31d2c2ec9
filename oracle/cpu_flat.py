"""CPU timing twin of the reference's exact search (TEST/BENCH INFRASTRUCTURE ONLY).

Restates what faiss_cpu==1.10.0's IndexFlatIP.search does (faiss/utils/distances.cpp, not vendored
in /root/reference; call site main/search_from_image.py:247):

* nq >= 20 — `exhaustive_inner_product_blas`: query blocks of 4096 x database blocks of 1024 rows,
  one fp32 `sgemm` per block pair (here numpy -> multi-threaded OpenBLAS), then the block's scores
  are pushed into one k-entry min-heap per query by an OpenMP loop over the queries
  (oracle/cpu_flat.c: flat_heap_addn), and the heaps are reordered at the end;
* nq < 20 — `exhaustive_inner_product_seq`: an OpenMP loop over the queries, each scanning all rows
  with a SIMD dot product (flat_scan_seq).

faiss itself is not installable here (no wheel, no network), so bench.py labels numbers from this
module kind="port".  Ordering of exact fp32 near-ties follows BLAS summation order, as in faiss;
it is NOT the bit-exact Spec R oracle (b2k_oracle.c) and is never used for parity.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "libcpu_flat.so"
_lib = None

BS_X, BS_Y = 4096, 1024      # faiss: distance_compute_blas_query_bs / distance_compute_blas_database_bs


def build(force: bool = False) -> Path:
    src = _HERE / "cpu_flat.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libcpu_flat.so"], check=True, capture_output=True)
    return _LIB


def _L():
    global _lib
    if _lib is None:
        build()
        # the OpenMP loops alternate with OpenBLAS's pthread pool: idle OpenMP workers must sleep, not spin
        # on the cores the sgemm needs (read by libgomp when it is first loaded)
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        os.environ.setdefault("GOMP_SPINCOUNT", "0")
        lib = C.CDLL(str(_LIB))
        lib.flat_heap_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        lib.flat_heap_addn.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                       C.c_void_p, C.c_void_p]
        lib.flat_scan_seq.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p]
        lib.flat_heap_reorder.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.flat_num_threads.restype = C.c_int32
        lib.flat_set_num_threads.argtypes = [C.c_int32]
        for f in (lib.flat_heap_init, lib.flat_heap_addn, lib.flat_scan_seq, lib.flat_heap_reorder,
                  lib.flat_set_num_threads):
            f.restype = None
        _lib = lib
    return _lib


def set_num_threads(n: int) -> None:
    """OpenMP threads of the heap / scan loops (the sgemm's threads are OpenBLAS's: threadpoolctl)."""
    _L().flat_set_num_threads(int(n))


def num_threads() -> int:
    return int(_L().flat_num_threads())


def search_flat_ip(db: np.ndarray, q: np.ndarray, k: int):
    """(ip [nq,k] descending, labels [nq,k]) with (-inf, -1) padding when k > ntotal."""
    lib = _L()
    db = np.ascontiguousarray(db, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = db.shape
    nq = q.shape[0]
    hv = np.empty((nq, k), np.float32)
    hi = np.empty((nq, k), np.int64)
    lib.flat_heap_init(hv.ctypes.data, hi.ctypes.data, nq * k)
    if nq < 20:
        lib.flat_scan_seq(db.ctypes.data, n, d, q.ctypes.data, nq, k, hv.ctypes.data, hi.ctypes.data)
    else:
        ip = np.empty((min(nq, BS_X), BS_Y), np.float32)
        for i0 in range(0, nq, BS_X):
            i1 = min(nq, i0 + BS_X)
            for j0 in range(0, n, BS_Y):
                j1 = min(n, j0 + BS_Y)
                blk = ip[: i1 - i0, : j1 - j0]
                np.matmul(q[i0:i1], db[j0:j1].T, out=blk)                      # sgemm
                lib.flat_heap_addn(blk.ctypes.data, i1 - i0, j1 - j0, ip.shape[1], j0, k,
                                   hv[i0:i1].ctypes.data, hi[i0:i1].ctypes.data)
    lib.flat_heap_reorder(nq, k, hv.ctypes.data, hi.ctypes.data)
    return hv, hi
