"""N > 1 host logic on CPU: world_size-2 gloo run of the row-sharded search plumbing
(shard ranges, packed all-gather, unpack, merge order).  The local scorer and the merge are
injected CPU stand-ins backed by the oracle; on a GPU box the same class drives FlatShard and
the K-merge kernel (bench.py, tests/test_gpu_parity.py::test_base_offset_and_two_shard_merge)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import DIMS
from image_recommender_b200.sharded import ShardedSearcher, shard_range


def test_shard_ranges_partition_rows():
    for n in (0, 1, 7, 8, 9, 1000, 10_000_000):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert all(0 <= b - a <= -(-n // w) for a, b in rs)
    assert shard_range(10_000_000, 8, 3) == (3_750_000, 5_000_000)


def _worker(rank, world, port, n, nq, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
        q = oracle.synth_queries(DIMS, nq, n, n_clusters=4)
        r0, r1 = shard_range(n, world, rank)

        def local_search(qt, kk, out):
            d, l, ip = oracle.search_exact(pk["f32"][r0:r1], qt.numpy(), kk, pk["norm2"][r0:r1], base_offset=r0)
            out[0].copy_(torch.from_numpy(d)); out[1].copy_(torch.from_numpy(l)); out[2].copy_(torch.from_numpy(ip))

        def merge(ip, d, l):
            md, ml, mip = oracle.merge_topk(ip.numpy(), d.numpy(), l.numpy())
            return torch.from_numpy(md), torch.from_numpy(ml), torch.from_numpy(mip)

        s = ShardedSearcher(local_search, merge)
        assert s.world == world
        for rep in range(2):                    # second call reuses the exchange buffers
            d, l, ip = s.search_device(torch.from_numpy(q), k)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), d=d.numpy(), l=l.numpy(), ip=ip.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_matches_single_shard(tmp_path):
    n, nq, k, world = 1501, 6, 10, 2          # ragged: shards of 751 and 750 rows
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, nq, k, str(tmp_path)), nprocs=world, join=True)
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=4)
    wd, wl, wip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(got["l"], wl)
        assert np.array_equal(got["ip"].view(np.uint32), wip.view(np.uint32))
        assert np.array_equal(got["d"].view(np.uint32), wd.view(np.uint32))


def test_failed_local_search_publishes_its_epoch_before_raising():
    """A rank whose local search raises still owes its peers an epoch: ShardedSearcher calls exchange.skip() (the
    peers' merge kernels would otherwise wait out the 10 s timeout) and re-raises the original error."""
    import torch
    from image_recommender_b200.sharded import ShardedSearcher

    class FakeExchange:
        max_entries = 1 << 20
        skipped = 0

        def skip(self):
            self.skipped += 1

    def boom(q, k, out):
        raise RuntimeError("local search failed")

    ex = FakeExchange()
    s = ShardedSearcher(boom, lambda ip, d, l: (d[0], l[0], ip[0]), exchange=ex)
    s.world = 2                                  # as under torchrun with two ranks
    with pytest.raises(RuntimeError, match="local search failed"):
        s.search_device(torch.zeros((3, 8)), 5)
    assert ex.skipped == 1
    s.world = 1                                  # a single rank has nobody to tell
    with pytest.raises(RuntimeError):
        s.search_device(torch.zeros((3, 8)), 5)
    assert ex.skipped == 1
