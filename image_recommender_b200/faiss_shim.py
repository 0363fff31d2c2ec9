"""The slice of the `faiss` module namespace that the reference's retrieval path uses, served by
the B200 engine — so the *unmodified* reference files run on it with

    import image_recommender_b200.faiss_shim as faiss

Members and the call sites they stand in for (paths into the reference repository):
  IndexHNSWFlat(dim, M), .hnsw.efConstruction/.efSearch   main/create_index.py:219-221, 230-232
  IndexIVFPQ(quantizer, dim, nlist, m, nbits)             main/create_index.py:226
  index.is_trained / train / add / ntotal / search        :296-298, :311, :321; search_from_image.py:247, 340
  write_index / read_index                                main/create_index.py:320; search_from_image.py:339
  normalize_L2                                            main/search_from_image.py:322

Exact search replaces the approximate indexes; M / efConstruction / efSearch / nlist / m / nbits are
accepted and ignored.  `index.table_dims = [48, 128, 1792]` (before the first add) tells the engine
where the per-table parts of a concatenated row are, so that each part is L2-normalised; without it
the whole row is treated as one table.  GPU-only: every call raises B2KError without a B200.
"""
from __future__ import annotations

import numpy as np

from .index import FlatShard, file_info, normalize_L2  # noqa: F401  (normalize_L2 re-exported)


class _HnswParams:
    efConstruction = 40
    efSearch = 16


class _ExactIndex:
    def __init__(self, d: int, *ignored, device: int = 0):
        self.d = int(d)
        self.device = device
        self.table_dims = None
        self.hnsw = _HnswParams()
        self.is_trained = True
        self._shard: FlatShard | None = None

    def _ensure(self, capacity: int) -> FlatShard:
        if self._shard is None:
            dims = list(self.table_dims) if self.table_dims else [self.d]
            if sum(dims) != self.d:
                raise ValueError(f"table_dims {dims} do not add up to d={self.d}")
            self._shard = FlatShard(dims, capacity, device=self.device)
        return self._shard

    @property
    def ntotal(self) -> int:
        return 0 if self._shard is None else self._shard.ntotal

    def train(self, x) -> None:
        return None

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        self._ensure(max(x.shape[0], 1024)).add(x)

    def search(self, x, k: int):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self._shard is None:
            return (np.full((x.shape[0], k), 3.4028235e38, np.float32), np.full((x.shape[0], k), -1, np.int64))
        return self._shard.search(x, k)


class IndexHNSWFlat(_ExactIndex):
    def __init__(self, d, M=32, metric=None, **kw):
        super().__init__(d, **kw)


class IndexFlatL2(_ExactIndex):
    pass


class IndexIVFPQ(_ExactIndex):
    def __init__(self, quantizer, d, nlist=0, m=0, nbits=0, **kw):
        super().__init__(d, **kw)


def write_index(index: _ExactIndex, path: str) -> None:
    index._ensure(1).save(str(path))


def read_index(path: str, device: int = 0) -> _ExactIndex:
    info = file_info(str(path))
    ix = _ExactIndex(sum(info["table_dims"]), device=device)
    ix.table_dims = info["table_dims"]
    ix._shard = FlatShard.load(str(path), device=device)
    return ix
