"""`FlatShard`: one GPU-resident exact-search shard behind the faiss index protocol.

Mirrors the members of the faiss index objects the reference touches
(main/create_index.py:219-234, 296-321; main/search_from_image.py:247, 339-340):
`add`, `search`, `ntotal`, `d`, `is_trained`, `train`.  All arithmetic happens in libb2k.so
(hand-written sm_100a CUDA); numpy/torch only carry buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _capi
from ._capi import B2KError, Stats, Synth, check

_lib = _capi.load_library()


def device_count() -> int:
    n = C.c_int32(0)
    st = _lib.b2k_device_count(C.byref(n))
    return int(n.value) if st == 0 else 0


def normalize_L2(x: np.ndarray, device: int = 0) -> None:
    """faiss.normalize_L2 (search_from_image.py:322): in place, zero rows untouched."""
    if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2 and x.flags.c_contiguous):
        raise TypeError("normalize_L2 expects a C-contiguous float32 [n, d] array (as faiss does)")
    check(_lib.b2k_normalize_l2(x.ctypes.data, x.shape[0], x.shape[1], device))


def _as_f32_2d(a, name: str) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.ndim != 2:
        raise ValueError(f"{name}: expected a 2-D float32 array")
    return a


class FlatShard:
    """Exact inner-product / squared-L2 top-k over rows [base_offset, base_offset + ntotal)."""

    is_trained = True   # create_index.py:296 checks this before train()

    def __init__(self, table_dims: Sequence[int], capacity: int, device: int = 0, base_offset: int = 0,
                 _handle: int | None = None):
        self.table_dims = [int(d) for d in table_dims]
        self.device = int(device)
        if _handle is None:
            dims = (C.c_int32 * len(self.table_dims))(*self.table_dims)
            h = C.c_void_p()
            check(_lib.b2k_create(dims, len(self.table_dims), int(capacity), self.device,
                                  int(base_offset), C.byref(h)))
            _handle = h.value
        self._h = C.c_void_p(_handle)

    # ---- lifetime ------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.b2k_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- faiss protocol ------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(_lib.b2k_ntotal(self._h))

    @property
    def d(self) -> int:
        return int(_lib.b2k_dim(self._h))

    @property
    def d_padded(self) -> int:
        return int(_lib.b2k_dim_padded(self._h))

    @property
    def base_offset(self) -> int:
        return int(_lib.b2k_base_offset(self._h))

    @property
    def capacity(self) -> int:
        return int(_lib.b2k_capacity(self._h))

    def train(self, x=None) -> None:   # create_index.py:298 — nothing to train for exact search
        return None

    def reset(self) -> None:
        """index.reset(): drop every row, keep the allocation."""
        check(_lib.b2k_reset(self._h))

    def reserve(self, capacity: int) -> None:
        check(_lib.b2k_reserve(self._h, int(capacity)))

    def _split(self, x: np.ndarray) -> list[np.ndarray]:
        if x.shape[1] != self.d:
            raise ValueError(f"add: rows have {x.shape[1]} columns, index dimension is {self.d}")
        out, off = [], 0
        for d in self.table_dims:
            out.append(np.ascontiguousarray(x[:, off:off + d]))
            off += d
        return out

    def add(self, x) -> None:
        """index.add(arr) (create_index.py:311): arr = concatenated rows [n, D] float32."""
        x = _as_f32_2d(x, "add")
        self.add_tables(self._split(x))

    def add_tables(self, tables: Sequence[np.ndarray]) -> None:
        """Same with the per-table arrays [n, d_t] not yet concatenated (skips a host copy)."""
        tabs = [_as_f32_2d(t, "add_tables") for t in tables]
        if len(tabs) != len(self.table_dims) or any(t.shape[1] != d for t, d in zip(tabs, self.table_dims)):
            raise ValueError("add_tables: table shapes do not match the index")
        n = tabs[0].shape[0]
        if any(t.shape[0] != n for t in tabs):
            raise ValueError("add_tables: tables disagree on the row count")
        if self.ntotal + n > self.capacity:
            # amortised doubling while it fits; a shard that already fills most of the GPU grows to exactly what
            # is needed instead (the re-allocation holds the old and the new arrays at once)
            try:
                self.reserve(max(self.ntotal + n, 2 * self.capacity, 1024))
            except B2KError:
                self.reserve(self.ntotal + n)
        ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        check(_lib.b2k_add(self._h, ptrs, n))

    def ingest_sqlite(self, db_path: str, sql: str, max_rows: int, rows_per_slot: int = 4096) -> np.ndarray:
        """Native build loop (csrc/ingest.cu): run `sql` (id, blob_1..blob_T) through libsqlite3, view
        the float32 payload of every blob in place, stage rows in pinned memory and K-pack them while
        the next slot is decoded.  Returns the image ids of the appended rows, in offset order.
        Raises B2KError(status=E_UNSUPPORTED) on the first blob the strict recogniser does not know
        (rows committed before it stay appended: reset() and decode in Python instead)."""
        ids = np.empty((max(int(max_rows), 1),), np.int64)
        n = C.c_int64(0)
        check(_lib.b2k_stage_open(self._h, int(rows_per_slot)))
        try:
            check(_lib.b2k_ingest_sqlite(self._h, str(db_path).encode(), sql.encode(), ids.ctypes.data,
                                         int(max_rows), C.byref(n)))
        finally:
            _lib.b2k_stage_close(self._h)
        return ids[:n.value].copy()

    def ingest_sqlite_mt(self, db_path: str, sql_range: str, id_bounds, max_rows: int, n_threads: int,
                         rows_per_slot: int = 4096) -> np.ndarray:
        """`ingest_sqlite` with `n_threads` reader threads (b2k_ingest_sqlite_mt): `sql_range` is the query with
        `i.id >= ?1 AND i.id < ?2`, `id_bounds` cuts the ids into chunks of at most `rows_per_slot` images.  Every
        thread has its own connection and two staging slots; chunks are committed in order, so rows, offsets and
        the returned ids equal those of the single-threaded call."""
        bounds = np.ascontiguousarray(id_bounds, dtype=np.int64)
        n_chunks = int(bounds.size) - 1
        n_threads = max(1, min(int(n_threads), 16, max(n_chunks, 1)))
        ids = np.empty((max(int(max_rows), 1),), np.int64)
        n = C.c_int64(0)
        check(_lib.b2k_stage_open_n(self._h, int(rows_per_slot), 2 * n_threads))
        try:
            check(_lib.b2k_ingest_sqlite_mt(self._h, str(db_path).encode(), sql_range.encode(), bounds.ctypes.data,
                                            n_chunks, n_threads, ids.ctypes.data, int(max_rows), C.byref(n)))
        finally:
            _lib.b2k_stage_close(self._h)
        return ids[:n.value].copy()

    def search(self, q, k: int):
        """index.search(query_vec, k) (search_from_image.py:247) -> (distances, labels)."""
        dist, lab, _ = self.search_ip(q, k, want_ip=False)
        return dist, lab

    def search_ip(self, q, k: int, want_ip: bool = True):
        q = _as_f32_2d(q, "search")
        if q.shape[1] != self.d:
            raise ValueError(f"search: queries have {q.shape[1]} columns, index dimension is {self.d}")
        nq = q.shape[0]
        dist = np.empty((nq, k), np.float32)
        lab = np.empty((nq, k), np.int64)
        ip = np.empty((nq, k), np.float32) if want_ip else None
        if nq:
            check(_lib.b2k_search(self._h, q.ctypes.data, nq, int(k), dist.ctypes.data, lab.ctypes.data,
                                  ip.ctypes.data if want_ip else None))
        return dist, lab, ip

    def search_groups(self, parts, group_offsets, k: int, want_ip: bool = False):
        """One query per GROUP of image vectors (search_from_image.py:305-322 + :247): parts [n_images, D] float32,
        group g = rows [group_offsets[g], group_offsets[g+1]).  Mean, whole-vector normalisation and search all
        happen on the device (b2k_search_groups); returns (distances, labels[, ip]) with one row per group."""
        parts = _as_f32_2d(parts, "search_groups")
        offs = np.ascontiguousarray(group_offsets, dtype=np.int32)
        if parts.shape[1] != self.d:
            raise ValueError(f"search_groups: vectors have {parts.shape[1]} columns, index dimension is {self.d}")
        ng = offs.size - 1
        dist = np.empty((ng, k), np.float32)
        lab = np.empty((ng, k), np.int64)
        ip = np.empty((ng, k), np.float32) if want_ip else None
        if ng > 0:
            check(_lib.b2k_search_groups(self._h, parts.ctypes.data, parts.shape[0], offs.ctypes.data, ng, int(k),
                                         dist.ctypes.data, lab.ctypes.data, ip.ctypes.data if want_ip else None))
        return (dist, lab, ip) if want_ip else (dist, lab)

    # ---- device-resident (torch) variants -------------------------------------------------
    def search_device(self, q, k: int, out=None, stream=None):
        """q: CUDA float32 [nq, D] torch tensor on this shard's device.  Enqueues on the current
        torch stream and returns (dist, labels, ip) CUDA tensors without synchronising."""
        import torch
        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.dim() == 2
        assert q.device.index == self.device and q.shape[1] == self.d
        nq = q.shape[0]
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32, device=q.device),
                   torch.empty((nq, k), dtype=torch.int64, device=q.device),
                   torch.empty((nq, k), dtype=torch.float32, device=q.device))
        dist, lab, ip = out
        st = torch.cuda.current_stream(q.device).cuda_stream if stream is None else stream
        check(_lib.b2k_search_device(self._h, q.data_ptr(), nq, int(k), dist.data_ptr(), lab.data_ptr(),
                                     ip.data_ptr(), C.c_void_p(st)))
        return dist, lab, ip

    def add_tables_device(self, tables, stream=None) -> None:
        import torch
        n = tables[0].shape[0]
        for t, d in zip(tables, self.table_dims):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (n, d)
        st = torch.cuda.current_stream(tables[0].device).cuda_stream if stream is None else stream
        ptrs = (C.c_void_p * len(tables))(*[t.data_ptr() for t in tables])
        check(_lib.b2k_add_device(self._h, ptrs, n, C.c_void_p(st)))

    # ---- options / introspection ----------------------------------------------------------
    def set_option(self, key: int, value: int) -> None:
        check(_lib.b2k_set_option(self._h, int(key), int(value)))

    def stats(self) -> dict:
        s = Stats()
        check(_lib.b2k_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in Stats._fields_}

    def get_rows(self, row0: int, n: int):
        """(fp32 rows [n, D], bf16 bit patterns [n, Dp], ||row||² [n]) as packed on the device."""
        f = np.empty((n, self.d), np.float32)
        b = np.empty((n, self.d_padded), np.uint16)
        n2 = np.empty((n,), np.float32)
        check(_lib.b2k_get_rows(self._h, int(row0), int(n), f.ctypes.data, b.ctypes.data, n2.ctypes.data))
        return f, b, n2

    # ---- persistence (faiss.write_index / read_index) ---------------------------------------
    def save(self, path: str, ids=None) -> None:
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            check(_lib.b2k_save(self._h, str(path).encode(), ids.ctypes.data, ids.size))
        else:
            check(_lib.b2k_save(self._h, str(path).encode(), None, 0))

    def save_shard(self, path: str, ids, file_row_begin: int, file_total_rows: int, create: bool) -> None:
        """Row-sharded build: write this shard's rows (and ids) at their place in one shared index file."""
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            if ids.size != self.ntotal:
                raise ValueError(f"save_shard: {ids.size} ids for {self.ntotal} rows")
        check(_lib.b2k_save_shard(self._h, str(path).encode(), ids.ctypes.data if ids is not None else None,
                                  int(file_row_begin), int(file_total_rows), 1 if create else 0))

    @classmethod
    def load(cls, path: str, device: int = 0, row_begin: int = 0, row_end: int = -1, capacity: int = 0) -> "FlatShard":
        """capacity: rows to allocate for (an index about to grow is loaded into its final size: no realloc)."""
        info = file_info(path)
        h = C.c_void_p()
        check(_lib.b2k_load(str(path).encode(), int(device), int(row_begin), int(row_end), int(capacity), C.byref(h)))
        return cls(info["table_dims"], 0, device=device, _handle=h.value)

    # ---- synthetic data (bench / tests) -------------------------------------------------------
    @staticmethod
    def _synth(seed, n_clusters, sigma, abs_mask, total_rows) -> Synth:
        return Synth(seed=seed, n_clusters=n_clusters, sigma=sigma, abs_mask=abs_mask, total_rows=total_rows)

    def fill_synthetic(self, n: int, total_rows: int, seed: int = 0xC0FFEE, n_clusters: int = 4096,
                       sigma: float = 0.3, abs_mask: int = 1) -> None:
        p = self._synth(seed, n_clusters, sigma, abs_mask, total_rows)
        check(_lib.b2k_fill_synthetic(self._h, int(n), C.byref(p)))

    def synth_queries_device(self, nq: int, total_rows: int, seed: int = 0xC0FFEE, n_clusters: int = 4096,
                             sigma: float = 0.3, abs_mask: int = 1, qseed: int = 0x5EED,
                             sigma_q: float = 0.05):
        import torch
        q = torch.empty((nq, self.d), dtype=torch.float32, device=f"cuda:{self.device}")
        p = self._synth(seed, n_clusters, sigma, abs_mask, total_rows)
        st = torch.cuda.current_stream(q.device).cuda_stream
        check(_lib.b2k_synth_queries_device(self._h, nq, C.byref(p), qseed, sigma_q, q.data_ptr(),
                                            C.c_void_p(st)))
        return q


def parse_f32_blob(blob: bytes):
    """The strict blob recogniser of the native ingest (host only): float32 vector viewed inside a
    pickled 1-D ndarray, or None when the blob is in any other format."""
    ptr, d = C.c_void_p(), C.c_int64(0)
    if _lib.b2k_parse_f32_blob(blob, len(blob), C.byref(ptr), C.byref(d)) != 0:
        return None
    off = ptr.value - C.cast(C.c_char_p(blob), C.c_void_p).value
    return np.frombuffer(blob, dtype="<f4", count=d.value, offset=off)


def file_info(path: str) -> dict:
    n = C.c_int64(0)
    nt = C.c_int32(0)
    dims = (C.c_int32 * _capi.B2K_MAX_TABLES)()
    has = C.c_int32(0)
    check(_lib.b2k_file_info(str(path).encode(), C.byref(n), C.byref(nt), dims, C.byref(has)))
    return {"n_rows": int(n.value), "table_dims": [int(dims[i]) for i in range(nt.value)],
            "has_ids": bool(has.value)}


def load_ids(path: str, row_begin: int, n: int) -> np.ndarray:
    out = np.empty((n,), np.int64)
    check(_lib.b2k_load_ids(str(path).encode(), int(row_begin), int(n), out.ctypes.data))
    return out


def merge_topk_device(ip, dist, labels, out=None, stream=None):
    """[G, nq, k] CUDA tensors (per-shard results) -> merged (dist, labels, ip) [nq, k]."""
    import torch
    G, nq, k = ip.shape
    assert ip.is_contiguous() and dist.is_contiguous() and labels.is_contiguous()
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.float32, device=ip.device),
               torch.empty((nq, k), dtype=torch.int64, device=ip.device),
               torch.empty((nq, k), dtype=torch.float32, device=ip.device))
    o_d, o_l, o_ip = out
    st = torch.cuda.current_stream(ip.device).cuda_stream if stream is None else stream
    check(_lib.b2k_merge_topk_device(ip.data_ptr(), dist.data_ptr(), labels.data_ptr(), G, nq, k,
                                     o_ip.data_ptr(), o_d.data_ptr(), o_l.data_ptr(), ip.device.index,
                                     C.c_void_p(st)))
    return o_d, o_l, o_ip


__all__ = ["FlatShard", "B2KError", "normalize_L2", "device_count", "file_info", "load_ids",
           "merge_topk_device", "parse_f32_blob"]
