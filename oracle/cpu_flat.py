"""CPU timing twin of the reference's exact search (TEST/BENCH INFRASTRUCTURE ONLY).

Restates what faiss_cpu==1.10.0's IndexFlatIP.search does for nq >= 20 — fp32 `sgemm` over
blocks of 1024 database rows followed by a running top-k merge (faiss/utils/distances.cpp:
exhaustive_inner_product_blas, bs_y = 1024) — and for nq < 20 the per-query SIMD dot loop
(exhaustive_inner_product_seq), parallel over queries only.  faiss itself is not installable
here (no wheel, no network), so bench.py labels numbers from this module kind="port".
Ordering of exact fp32 near-ties follows BLAS summation order, as in faiss; it is NOT the
bit-exact Spec R oracle (b2k_oracle.c) and is never used for parity.
"""
from __future__ import annotations

import numpy as np


def search_flat_ip(db: np.ndarray, q: np.ndarray, k: int, block: int = 1024):
    """(ip [nq,k] descending, labels [nq,k]) with -1 padding when k > ntotal."""
    n, nq = db.shape[0], q.shape[0]
    best_s = np.full((nq, k), -np.inf, np.float32)
    best_i = np.full((nq, k), -1, np.int64)
    if nq < 20:
        # faiss: one query at a time against all rows (sequential kernel)
        for i in range(nq):
            s = db @ q[i]
            kk = min(k, n)
            idx = np.argpartition(-s, kk - 1)[:kk] if kk < n else np.arange(n)
            idx = idx[np.argsort(-s[idx], kind="stable")]
            best_s[i, :kk] = s[idx]
            best_i[i, :kk] = idx
        return best_s, best_i
    # blocks sized so that the nq x block score tile stays cache-resident, as faiss does
    step = block * max(1, 4096 // max(nq, 1)) if nq < 4096 else block
    for r0 in range(0, n, step):
        blk = db[r0:r0 + step]
        s = q @ blk.T                                     # sgemm
        m = blk.shape[0]
        kk = min(k, m)
        part = np.argpartition(-s, kk - 1, axis=1)[:, :kk] if kk < m else np.tile(np.arange(m), (nq, 1))
        ps = np.take_along_axis(s, part, axis=1)
        cat_s = np.concatenate([best_s, ps], axis=1)
        cat_i = np.concatenate([best_i, part + r0], axis=1)
        sel = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]
        best_s = np.take_along_axis(cat_s, sel, axis=1)
        best_i = np.take_along_axis(cat_i, sel, axis=1)
    return best_s, best_i
