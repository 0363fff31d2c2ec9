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


class ShardedSearcher:
    """Wraps the local shard's `search_device`; `group=None` or world size 1 means no exchange.

    `local_search(q, k, out)` and `merge(ip, dist, labels)` are injectable so that the
    collective plumbing can be exercised with the gloo backend on CPU tensors (tests/)."""

    def __init__(self, local_search: Callable, merge: Callable, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.local_search = local_search
        self.merge = merge
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._send = None
        self._recv = None

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
        self.local_search(q, k, (dist_t, lab, ip_t))
        if self.world == 1:
            return dist_t, lab, ip_t
        self._dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        g_lab = recv[:, :m].contiguous().view(self.world, nq, k)
        g_fl = recv[:, m:].contiguous().view(torch.float32).view(self.world, 2, nq, k)
        g_dist = g_fl[:, 0].contiguous()
        g_ip = g_fl[:, 1].contiguous()
        return self.merge(g_ip, g_dist, g_lab)
