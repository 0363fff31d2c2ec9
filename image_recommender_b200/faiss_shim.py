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

Distances.  The engine stores L2-normalised parts and ranks by inner product, which equals the reference's
METRIC_L2 ranking exactly when all rows share one norm (the extractors emit unit-norm parts).
Without `table_dims` the whole row is normalised: rows of norm sqrt(T) come back with distances 2 - 2cos
instead of 1 + T - 2ip (same order, another scale; a warning says so once).  Rows whose norms DIFFER would be
ranked differently from an L2 search over the raw rows: `add` refuses them (ValueError) unless
`index.allow_renormalize = True` — silently altering the ranking is not an option.
"""
from __future__ import annotations

import logging

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
        self.allow_renormalize = False
        self._warned_scale = False
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
        if not self.table_dims and x.shape[0]:
            # whole-row normalisation: harmless for equal-norm rows (order kept, distances rescaled), a different
            # ranking from the reference's L2 index otherwise
            norms = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float64))
            lo, hi = float(norms.min()), float(norms.max())
            if hi > 0 and (hi - lo) > 1e-3 * hi and not self.allow_renormalize:
                raise ValueError(
                    f"faiss_shim: rows have norms {lo:.4g} .. {hi:.4g}; the engine would normalise them and rank by "
                    "cosine, which differs from the reference's L2 search over the raw rows.  Set index.table_dims "
                    "(per-table unit-norm parts) or index.allow_renormalize = True to accept cosine ranking.")
            if abs(hi - 1.0) > 1e-3 and not self._warned_scale:
                logging.warning("faiss_shim: rows of norm %.4g are stored normalised (no table_dims): returned distances "
                                "are 2 - 2cos instead of the raw squared L2 (same order)", hi)
                self._warned_scale = True
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
