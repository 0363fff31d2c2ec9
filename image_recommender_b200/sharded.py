"""Row-sharded search across the GPUs of one box: one process per GPU, each owns a
contiguous offset range; queries are replicated; per-shard top-k lists are exchanged with ONE
all-gather (B·k records of 16 B per rank, latency-bound) and merged by the K-merge kernel with
the same total order on every rank (higher ip, then lower offset).
"""
from __future__ import annotations

from typing import Callable


def shard_range(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Rows [begin, end) of shard `rank`: contiguous ranges of ceil(n/world) rows, so that
    global offset = begin + local row keeps the reference's offset -> image_id mapping
    (create_index.py:236-249) trivial."""
    per = -(-n_total // world_size)
    begin = min(n_total, rank * per)
    return begin, min(n_total, begin + per)


class PeerExchange:
    """NVLink peer-memory exchange + merge (csrc/xchg.cu): every rank stores its top-k records into
    all ranks' receive buffers and merges locally.  Set up once per process group."""

    def __init__(self, device: int, rank: int, world: int, max_entries: int = 4096 * 32, group=None,
                 _local_peers=None):
        import ctypes as C
        from . import _capi
        self._C, self._capi = C, _capi
        self._lib = _capi.load_library()
        self.device, self.rank, self.world = device, rank, world
        self.max_entries = int(max_entries)
        h = C.c_void_p()
        _capi.check(self._lib.b2k_xchg_create(device, rank, world, max_entries, C.byref(h)))
        self._h = h
        self._handle = (C.c_ubyte * 64)()
        raw = C.c_void_p()
        _capi.check(self._lib.b2k_xchg_handle(self._h, self._handle, C.byref(raw)))
        self.raw_ptr = raw.value
        if world > 1 and _local_peers is None:
            import torch
            import torch.distributed as dist
            mine = torch.tensor(list(self._handle), dtype=torch.uint8, device=f"cuda:{device}")
            allh = torch.empty((world, 64), dtype=torch.uint8, device=f"cuda:{device}")
            dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
            buf = (C.c_ubyte * (64 * world)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
            _capi.check(self._lib.b2k_xchg_connect(self._h, buf, None))
            dist.barrier(group=group)     # every rank has mapped every buffer before the first push

    def connect_local(self, raw_ptrs) -> None:
        """Ranks living in one process (tests): plain device pointers instead of IPC handles."""
        C = self._C
        arr = (C.c_void_p * self.world)(*raw_ptrs)
        self._capi.check(self._lib.b2k_xchg_connect(self._h, None, arr))

    def push(self, ip, dist_t, lab, stream=None) -> None:
        import torch
        nq, k = ip.shape
        st = torch.cuda.current_stream(ip.device).cuda_stream if stream is None else stream
        self._capi.check(self._lib.b2k_xchg_push(self._h, ip.data_ptr(), dist_t.data_ptr(), lab.data_ptr(), nq, k,
                                                 self._C.c_void_p(st)))

    def skip(self, stream=None) -> None:
        """Publish this rank's next epoch without records (its local search failed; the peers must not wait)."""
        import torch
        st = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream if stream is None else stream
        self._capi.check(self._lib.b2k_xchg_skip(self._h, self._C.c_void_p(st)))

    def merge(self, nq: int, k: int, out=None, stream=None):
        import torch
        dev = torch.device("cuda", self.device)
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32, device=dev),
                   torch.empty((nq, k), dtype=torch.int64, device=dev),
                   torch.empty((nq, k), dtype=torch.float32, device=dev))
        o_d, o_l, o_ip = out
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        self._capi.check(self._lib.b2k_xchg_merge(self._h, nq, k, o_ip.data_ptr(), o_d.data_ptr(), o_l.data_ptr(),
                                                  self._C.c_void_p(st)))
        return o_d, o_l, o_ip

    def check(self) -> None:
        """Synchronises the device; raises B2KError(E_PEER) if a merge since the last call gave up waiting for a
        rank (its outputs were padded with label -1)."""
        bits = self._C.c_uint32(0)
        self._capi.check(self._lib.b2k_xchg_status(self._h, self._C.byref(bits)))

    def close(self) -> None:
        if self._h:
            self._lib.b2k_xchg_destroy(self._h)
            self._h = None


class ShardedSearcher:
    """Wraps the local shard's `search_device`; `group=None` or world size 1 means no exchange.

    `local_search(q, k, out)` and `merge(ip, dist, labels)` are injectable so that the
    collective plumbing can be exercised with the gloo backend on CPU tensors (tests/)."""

    def __init__(self, local_search: Callable, merge: Callable, group=None, exchange: "PeerExchange | None" = None):
        import torch.distributed as dist
        self._dist = dist
        self.local_search = local_search
        self.merge = merge
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._send = None
        self._recv = None
        self.exchange = exchange      # NVLink peer-memory path instead of all-gather + merge
        self._out = None

    def _buffers(self, nq: int, k: int, device):
        import torch
        need = 2 * nq * k
        if self._send is None or self._send.numel() != need or self._send.device != device:
            self._send = torch.empty(need, dtype=torch.int64, device=device)
            self._recv = torch.empty((self.world, need), dtype=torch.int64, device=device)
        return self._send, self._recv

    def search_device(self, q, k: int):
        """q: [nq, D] float32 on this rank's device, identical on every rank.
        Returns (dist [nq,k] f32, labels [nq,k] i64, ip [nq,k] f32), identical on every rank."""
        import torch
        nq = q.shape[0]
        send, recv = self._buffers(nq, k, q.device)
        m = nq * k
        lab = send[:m].view(nq, k)
        fl = send[m:].view(torch.float32)            # 2*m floats
        dist_t = fl[:m].view(nq, k)
        ip_t = fl[m:].view(nq, k)
        try:
            self.local_search(q, k, (dist_t, lab, ip_t))
        except Exception:
            # this rank still owes its peers an epoch: without it every peer's merge waits out the 10 s timeout
            if self.world > 1 and self.exchange is not None and nq * k <= self.exchange.max_entries:
                self.exchange.skip()
            raise
        if self.world == 1:
            return dist_t, lab, ip_t
        if self.exchange is not None and nq * k <= self.exchange.max_entries:
            if self._out is None or self._out[0].shape != (nq, k):
                self._out = (torch.empty((nq, k), dtype=torch.float32, device=q.device),
                             torch.empty((nq, k), dtype=torch.int64, device=q.device),
                             torch.empty((nq, k), dtype=torch.float32, device=q.device))
            self.exchange.push(ip_t, dist_t, lab)
            return self.exchange.merge(nq, k, out=self._out)
        self._dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        g_lab = recv[:, :m].contiguous().view(self.world, nq, k)
        g_fl = recv[:, m:].contiguous().view(torch.float32).view(self.world, 2, nq, k)
        g_dist = g_fl[:, 0].contiguous()
        g_ip = g_fl[:, 1].contiguous()
        return self.merge(g_ip, g_dist, g_lab)


class ShardedIndex:
    """faiss index protocol (`search`, `ntotal`, `d`) over a row-sharded index file: every rank of
    the process group loads ITS row range of the file (b2k_load's range form), searches it locally
    and exchanges the per-shard top-k (K-exchange over NVLink peer memory, or NCCL all-gather +
    K-merge).  SPMD: every rank calls `search` with the same queries and gets the same global
    result — ids are the file's global offsets, exactly what a single-GPU load returns."""

    def __init__(self, shard, n_total: int, group=None, exchange: "PeerExchange | None" = None):
        from .index import merge_topk_device
        self.shard = shard
        self.n_total = int(n_total)
        self.exchange = exchange
        self._searcher = ShardedSearcher(lambda q, k, out: shard.search_device(q, k, out=out),
                                         lambda ip, d, l: merge_topk_device(ip, d, l),
                                         group=group, exchange=exchange)

    @classmethod
    def load(cls, path, device: int, group=None, peer: bool = True, max_batch: int = 4096) -> "ShardedIndex":
        import torch.distributed as dist
        from .index import FlatShard, file_info
        info = file_info(path)
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        r0, r1 = shard_range(info["n_rows"], world, rank)
        shard = FlatShard.load(path, device=device, row_begin=r0, row_end=r1)
        exchange = None
        if world > 1 and peer:
            from ._capi import B2K_LIST
            # sized for the tuned case k <= 32; larger nq * k falls back to the all-gather path
            exchange = PeerExchange(device, rank, world, max_entries=max_batch * B2K_LIST, group=group)
        out = cls(shard, info["n_rows"], group=group, exchange=exchange)
        out._max_batch = max_batch
        return out

    @property
    def ntotal(self) -> int:
        return self.n_total

    @property
    def d(self) -> int:
        return self.shard.d

    is_trained = True

    def search(self, q, k: int):
        """q: float32 [nq, D] host array, identical on every rank -> (distances [nq,k], labels [nq,k])."""
        import numpy as np
        import torch
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        dev = torch.device("cuda", self.shard.device)
        dist_out = np.empty((q.shape[0], k), np.float32)
        lab_out = np.empty((q.shape[0], k), np.int64)
        step = getattr(self, "_max_batch", 4096)
        with torch.cuda.device(dev):
            for lo in range(0, q.shape[0], step):
                qd = torch.from_numpy(q[lo:lo + step]).to(dev)
                d_, l_, _ = self._searcher.search_device(qd, k)
                dist_out[lo:lo + step] = d_.cpu().numpy()
                lab_out[lo:lo + step] = l_.cpu().numpy()
        if self.exchange is not None:
            self.exchange.check()       # a rank that never published -> B2KError instead of -1 labels
        return dist_out, lab_out

    def search_groups(self, parts, group_offsets, k: int):
        """FlatShard.search_groups over the row-sharded index: every rank copies the (identical) image vectors to
        its GPU, runs the mean + normalise kernel there and searches its shard; one exchange as in `search`."""
        import ctypes as C
        import numpy as np
        import torch
        from . import _capi
        parts = np.ascontiguousarray(parts, dtype=np.float32)
        offs = np.ascontiguousarray(group_offsets, dtype=np.int32)
        ng = offs.size - 1
        dev = torch.device("cuda", self.shard.device)
        dist_out = np.empty((ng, k), np.float32)
        lab_out = np.empty((ng, k), np.int64)
        step = getattr(self, "_max_batch", 4096)
        lib = _capi.load_library()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            for lo in range(0, ng, step):
                hi = min(ng, lo + step)
                r0, r1 = int(offs[lo]), int(offs[hi])
                pd = torch.from_numpy(parts[r0:r1]).to(dev)
                od = torch.from_numpy(offs[lo:hi + 1] - r0).to(dev)
                qd = torch.empty((hi - lo, self.d), dtype=torch.float32, device=dev)
                _capi.check(lib.b2k_prep_groups_device(pd.data_ptr(), od.data_ptr(), hi - lo, self.d, qd.data_ptr(),
                                                       self.shard.device, C.c_void_p(st)))
                d_, l_, _ = self._searcher.search_device(qd, k)
                dist_out[lo:hi] = d_.cpu().numpy()
                lab_out[lo:hi] = l_.cpu().numpy()
        if self.exchange is not None:
            self.exchange.check()
        return dist_out, lab_out

    def close(self) -> None:
        if self.exchange is not None:
            self.exchange.close()
            self.exchange = None
        self.shard.close()
