"""`ShardGroup`: the row-sharded index on several GPUs of one box, driven by ONE process.

The reference's CLI and `ImageRecommender` are a single process (main/search_from_image.py:430-441); this is the
faiss index protocol (`search`, `ntotal`, `d`) over `b2k_group` (csrc/group.cu): one worker thread per device
inside the library, results pushed to the first device over NVLink peer memory and merged there.  No torchrun,
no torch.distributed, no NCCL — `ImageRecommender(device="all")` and `python -m main.search_from_image --device all`
use it.  The torchrun form (sharded.py) remains for one-process-per-GPU deployments.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _capi
from ._capi import check
from .index import FlatShard, device_count

_lib = _capi.load_library()


class ShardGroup:
    is_trained = True

    def __init__(self, devices: Sequence[int] | None = None):
        devs = list(range(device_count())) if devices is None else [int(d) for d in devices]
        if not devs:
            raise _capi.B2KError(_capi.E_NODEVICE, "no CUDA device (this engine has no CPU path)")
        arr = (C.c_int32 * len(devs))(*devs)
        h = C.c_void_p()
        check(_lib.b2k_group_create(arr, len(devs), C.byref(h)))
        self._h = h
        self.devices = devs
        self._adopted = []          # FlatShards handed over with set_shards(): kept alive here

    @classmethod
    def load(cls, path, devices: Sequence[int] | None = None) -> "ShardGroup":
        """Every device loads its contiguous row range of the index file (faiss.read_index, search_from_image.py:339)."""
        g = cls(devices)
        check(_lib.b2k_group_load(g._h, str(path).encode()))
        return g

    def set_shards(self, shards: Sequence[FlatShard]) -> None:
        """Adopt one FlatShard per rank (rank r on self.devices[r], base offsets = the shards' row ranges)."""
        if len(shards) != len(self.devices):
            raise ValueError("one shard per device")
        for r, s in enumerate(shards):
            check(_lib.b2k_group_set_shard(self._h, r, s._h))
        self._adopted = list(shards)

    def shard(self, rank: int) -> FlatShard:
        """Borrowed view of rank's shard (options, stats)."""
        if self._adopted:
            return self._adopted[rank]
        h = _lib.b2k_group_shard(self._h, rank)
        s = FlatShard.__new__(FlatShard)
        s.table_dims, s.device, s._h = [], self.devices[rank], C.c_void_p(h)
        s.close = lambda: None          # owned by the group
        return s

    @property
    def ntotal(self) -> int:
        return int(_lib.b2k_group_ntotal(self._h))

    @property
    def d(self) -> int:
        return int(_lib.b2k_group_dim(self._h))

    def search(self, q, k: int):
        dist, lab, _ = self.search_ip(q, k, want_ip=False)
        return dist, lab

    def search_ip(self, q, k: int, want_ip: bool = True):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != self.d:
            raise ValueError(f"search: queries have {q.shape[1]} columns, index dimension is {self.d}")
        nq = q.shape[0]
        dist = np.empty((nq, k), np.float32)
        lab = np.empty((nq, k), np.int64)
        ip = np.empty((nq, k), np.float32) if want_ip else None
        if nq:
            check(_lib.b2k_group_search(self._h, q.ctypes.data, nq, int(k), dist.ctypes.data, lab.ctypes.data,
                                        ip.ctypes.data if want_ip else None))
        return dist, lab, ip

    def search_groups(self, parts, group_offsets, k: int):
        parts = np.ascontiguousarray(parts, dtype=np.float32)
        offs = np.ascontiguousarray(group_offsets, dtype=np.int32)
        ng = offs.size - 1
        dist = np.empty((ng, k), np.float32)
        lab = np.empty((ng, k), np.int64)
        if ng > 0:
            check(_lib.b2k_group_search_groups(self._h, parts.ctypes.data, parts.shape[0], offs.ctypes.data, ng, int(k),
                                               dist.ctypes.data, lab.ctypes.data, None))
        return dist, lab

    # ---- resident-query form (bench): H2D once, then timed runs -----------------------------------
    def put_queries(self, q: np.ndarray, k: int) -> None:
        q = np.ascontiguousarray(q, dtype=np.float32)
        check(_lib.b2k_group_put_queries(self._h, q.ctypes.data, q.shape[0], int(k)))

    def run(self, nq: int, k: int) -> None:
        check(_lib.b2k_group_run(self._h, int(nq), int(k)))

    def last_run_ms(self) -> float:
        ms = C.c_float(0)
        check(_lib.b2k_group_last_run_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def get_results(self, nq: int, k: int):
        dist = np.empty((nq, k), np.float32)
        lab = np.empty((nq, k), np.int64)
        ip = np.empty((nq, k), np.float32)
        check(_lib.b2k_group_get_results(self._h, dist.ctypes.data, lab.ctypes.data, ip.ctypes.data))
        return dist, lab, ip

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.b2k_group_destroy(self._h)
            self._h = C.c_void_p(None)
        for s in self._adopted:
            s.close()
        self._adopted = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
