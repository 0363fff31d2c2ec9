"""GPU tests of the round-2 additions: bounded grid barrier under SM contention, the timing ring of
b2k_get_stats, load-with-capacity / append without re-allocation.  `pytest -m gpu` on a B200."""
import time

import numpy as np
import pytest

import oracle
from conftest import DIMS

pytestmark = pytest.mark.gpu


def _mk(n, dims=DIMS, n_clusters=8, seed=0xC0FFEE):
    tabs = oracle.synth_rows(dims, n, total_rows=n, n_clusters=n_clusters, seed=seed)
    return tabs, oracle.pack(tabs)


def test_two_indexes_on_two_streams_stay_exact_and_bounded(gpu):
    """The in-kernel seeding's grid barrier assumes co-resident CTAs.  Two indexes searched at the same time on
    two streams break that assumption (each scoring grid wants every SM; the second grid trickles onto the SMs the
    first one frees): the barrier gives up after 200 us of wall time (and is poisoned for the CTAs that arrive
    later), results stay bit-equal and the latency stays within a small margin of the serial time (it was 84 ms
    PER BARRIER before the bound).  Phase 1: two equal small-batch searches.  Phase 2: a 1024-query search (pair
    kernel, 16 query-tile waves) with the small-batch seeded search launched into its tail."""
    import torch
    import image_recommender_b200 as irb
    n, nq, k = 1_000_000, 32, 10            # 3907 tiles / 148 splits = 26 per split: seeded, inside the scoring launch
    shards, qs, want = [], [], []
    for s in range(2):
        ix = irb.FlatShard(DIMS, n, device=gpu)
        ix.fill_synthetic(n, total_rows=n, seed=0xC0FFEE + s)
        q = ix.synth_queries_device(nq, total_rows=n, seed=0xC0FFEE + s)
        shards.append(ix)
        qs.append(q)
        want.append([t.clone() for t in ix.search_device(q, k)])      # alone: the reference result of this index
        assert ix.stats()["launches"] == 5                            # no separate sampling pass: barrier in use
    big_q = shards[0].synth_queries_device(1024, total_rows=n, seed=7)
    big_want = [t.clone() for t in shards[0].search_device(big_q, k)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=gpu) for _ in range(2)]

    def serial_time(jobs, reps=5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for ix, q in jobs:
                ix.search_device(q, k)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def concurrent(jobs, wants, reps=10):
        worst = 0.0
        for it in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            outs = []
            for (ix, q), st in zip(jobs, streams):
                with torch.cuda.stream(st):
                    outs.append(ix.search_device(q, k))
            torch.cuda.synchronize()
            if it:                           # the first use of a fresh stream costs 4..50 ms of driver set-up
                worst = max(worst, time.perf_counter() - t0)
            for got, ref in zip(outs, wants):
                assert torch.equal(got[1], ref[1])
                assert torch.equal(got[0].view(torch.int32), ref[0].view(torch.int32))
        return worst

    jobs = list(zip(shards, qs))
    serial = serial_time(jobs)
    worst = concurrent(jobs, want)
    # at most two barriers of 200 us per search, plus scheduling noise: far below the old 2 x 84 ms
    assert worst < serial + 2e-3, (worst, serial)
    jobs2 = [(shards[0], big_q), (shards[1], qs[1])]
    serial2 = serial_time(jobs2)
    worst2 = concurrent(jobs2, [big_want, want[1]])
    assert worst2 < serial2 + 2e-3, (worst2, serial2)
    for ix in shards:
        ix.close()


def test_stats_average_back_to_back_passes(gpu):
    """b2k_get_stats averages the kernel times of the passes since the previous call (event ring inside the
    ABI): timing a loop of searches needs no host sync between them."""
    import torch
    import image_recommender_b200 as irb
    n = 200_000
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.fill_synthetic(n, total_rows=n)
    q = ix.synth_queries_device(8, total_rows=n)
    ix.search_device(q, 10)
    assert ix.stats()["n_timed"] == 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ix.search_device(q, 10)
    e1.record()
    torch.cuda.synchronize()
    st = ix.stats()
    assert st["n_timed"] == 20
    per_step = e0.elapsed_time(e1) / 20
    assert 0 < st["score_ms"] <= per_step and st["score_ms"] + st["tail_ms"] <= per_step * 1.05
    for _ in range(100):                       # more passes than the ring holds: the most recent 64
        ix.search_device(q, 10)
    assert ix.stats()["n_timed"] == 64
    ix.close()


def test_load_with_capacity_then_append_without_realloc(gpu, tmp_path):
    """--update loads the file straight into its final capacity (a re-allocation holds the old and the
    new arrays at once); appended rows and searches equal a one-shot build."""
    import image_recommender_b200 as irb
    n, extra = 3000, 500
    tabs, pk = _mk(n + extra)
    a = irb.FlatShard(DIMS, n, device=gpu)
    a.add_tables([t[:n] for t in tabs])
    f = tmp_path / "index_hnsw_x.faiss"
    a.save(str(f), ids=np.arange(n))
    a.close()
    b = irb.FlatShard.load(str(f), device=gpu, capacity=n + extra)
    assert b.ntotal == n and b.capacity == n + extra
    b.add_tables([t[n:] for t in tabs])
    assert b.capacity == n + extra                   # no reserve() happened
    f32, bf, n2 = b.get_rows(0, n + extra)
    assert np.array_equal(f32.view(np.uint32), pk["f32"].view(np.uint32))
    assert np.array_equal(bf, pk["bf16"])
    q = oracle.synth_queries(DIMS, 9, n + extra, n_clusters=8)
    dist, lab, ip = b.search_ip(q, 10)
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
    assert np.array_equal(lab, w_lab) and np.array_equal(ip.view(np.uint32), w_ip.view(np.uint32))
    b.close()


def test_update_refuses_index_without_id_column(gpu, tmp_path):
    """An index file without the image-id column (faiss_shim.write_index) cannot tell which images it holds:
    build_index(update_index=True) refuses up front instead of re-adding every row."""
    import sqlite3
    import image_recommender_b200 as irb
    from main.create_index import FAISSIndexBuilderDB
    tabs, _ = _mk(50, [48])
    ix = irb.FlatShard([48], 50, device=gpu)
    ix.add_tables(tabs)
    f = tmp_path / "index_hnsw_color.faiss"
    ix.save(str(f))                                   # no ids
    ix.close()
    db = tmp_path / "images.db"
    con = sqlite3.connect(db)
    con.executescript("CREATE TABLE images (id INTEGER PRIMARY KEY, path TEXT);"
                      "CREATE TABLE color_vectors (image_id INTEGER PRIMARY KEY, color_vector_blob BLOB);")
    con.commit()
    con.close()
    b = FAISSIndexBuilderDB(db_path=str(db), vector_types=["color"], index_file=str(f), log_dir=str(tmp_path))
    with pytest.raises(ValueError, match="no image-id column"):
        b.build_index(update_index=True)


@pytest.mark.parametrize("dims,n", [(DIMS, 4000), ([5, 3, 70], 1500)])
def test_search_groups_equals_host_mean_normalise_search(gpu, dims, n):
    """b2k_search_groups (mean over a group's image vectors + faiss.normalize_L2 + search, all on the device;
    search_from_image.py:305-322 + :247) == np.mean -> oracle normalise -> oracle search, bit for bit; the prep
    kernel alone (b2k_prep_groups_device) reproduces the host's query vectors."""
    import ctypes as C
    import torch
    import image_recommender_b200 as irb
    from image_recommender_b200 import _capi
    tabs, pk = _mk(n, dims)
    ix = irb.FlatShard(dims, n, device=gpu)
    ix.add_tables(tabs)
    d = sum(dims)
    rng = np.random.default_rng(5)
    sizes = [1, 2, 1, 5, 3, 1, 7, 2]
    imgs = oracle.synth_queries(dims, sum(sizes), n, n_clusters=8)
    imgs = (imgs * rng.uniform(0.5, 2.0, size=(imgs.shape[0], 1))).astype(np.float32)   # un-normalised inputs
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    want_q = []
    for g in range(len(sizes)):
        have = [imgs[i:i + 1] for i in range(offs[g], offs[g + 1])]
        want_q.append(np.ascontiguousarray(np.mean(have, axis=0), dtype=np.float32))    # the reference's expression
    want_q = oracle.normalize_l2(np.concatenate(want_q, axis=0))
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], want_q, 10, pk["norm2"])
    dist, lab, ip = ix.search_groups(imgs, offs, 10, want_ip=True)
    assert np.array_equal(lab, w_lab)
    assert np.array_equal(ip.view(np.uint32), w_ip.view(np.uint32))
    assert np.array_equal(dist.view(np.uint32), w_dist.view(np.uint32))
    pd, od = torch.from_numpy(imgs).cuda(gpu), torch.from_numpy(offs).cuda(gpu)
    qd = torch.empty((len(sizes), d), dtype=torch.float32, device=f"cuda:{gpu}")
    _capi.check(_capi.load_library().b2k_prep_groups_device(pd.data_ptr(), od.data_ptr(), len(sizes), d, qd.data_ptr(), gpu,
                                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert np.array_equal(qd.cpu().numpy().view(np.uint32), want_q.view(np.uint32))
    with pytest.raises(irb.B2KError):
        ix.search_groups(imgs, np.array([0, 3, 3, imgs.shape[0]], np.int32), 10)       # an empty group
    ix.close()


def test_fused_tail_equals_staged_tail(gpu):
    """B2K_OPT_FUSED_TAIL: the one-launch tail (cluster of CTAs per query: select -> re-rank -> finalize; K-collect
    and K-exact finishing their queries themselves) returns the bits of the one-launch-per-stage pipeline and of
    the oracle — plain queries, a near-duplicate burst (saturated lists -> K-collect), forced-exact queries and a
    candidate-budget overflow — and launches at most 5 kernels per search at batch 1."""
    import image_recommender_b200 as irb
    from image_recommender_b200 import _capi
    n = 30000
    tabs, _ = _mk(n)
    for t in tabs:                       # a run of 300 near-duplicates of row 1000, stored next to each other
        t[2000:2300] = t[1000] * (1.0 + 1e-4 * np.arange(300, dtype=np.float32)[:, None])
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.add_tables(tabs)
    q = oracle.synth_queries(DIMS, 700, n, n_clusters=8)
    q[5] = pk["f32"][2100] / np.sqrt(3.0)         # lands in the burst
    q[41] = pk["f32"][2299] / np.sqrt(3.0)
    for nq in (1, 7, 42, 130, 700):
        w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q[:nq], 10, pk["norm2"])
        res = {}
        for fused in (1, 0):
            ix.set_option(_capi.OPT_FUSED_TAIL, fused)
            res[fused] = ix.search_ip(q[:nq], 10)
            st = ix.stats()
            if fused and nq == 1:
                assert st["launches"] <= 5, st
            if nq >= 7:
                assert st["n_saturated"] > 0          # the burst saturates lists: K-collect ran
        for a, b in zip(res[1], res[0]):
            assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                  b.view(np.uint32) if b.dtype == np.float32 else b)
        assert np.array_equal(res[1][1], w_lab)
        assert np.array_equal(res[1][2].view(np.uint32), w_ip.view(np.uint32))
        assert np.array_equal(res[1][0].view(np.uint32), w_dist.view(np.uint32))
    ix.set_option(_capi.OPT_FUSED_TAIL, 1)
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q[:50], 10, pk["norm2"])
    for key, val in ((_capi.OPT_FORCE_EXACT, 1), (_capi.OPT_RERANK, 32), (_capi.OPT_COLLECT, 0)):
        ix.set_option(key, val)
        dist, lab, ip = ix.search_ip(q[:50], 10)
        assert ix.stats()["n_uncertified"] > 0
        assert np.array_equal(lab, w_lab) and np.array_equal(ip.view(np.uint32), w_ip.view(np.uint32))
        ix.set_option(key, 1 if key == _capi.OPT_COLLECT else 0)
    ix.close()


def _group_devices():
    """Two ranks: two GPUs when the box has them, else both on device 0 (the group logic is the same; the
    exchange then runs between two streams of one device)."""
    import image_recommender_b200 as irb
    return [0, 1] if irb.device_count() >= 2 else [0, 0]


def test_shard_group_single_process_equals_single_shard(gpu, tmp_path):
    """b2k_group (one process driving several row shards: worker thread per rank, peer-memory push to the root,
    merge there) returns exactly what ONE shard holding all rows returns — adopted shards and b2k_group_load of
    an index file alike, plain queries and query groups."""
    import image_recommender_b200 as irb
    n = 9001
    tabs, pk = _mk(n)
    whole = irb.FlatShard(DIMS, n, device=gpu)
    whole.add_tables(tabs)
    f = tmp_path / "index_hnsw_color_sift_dreamsim.faiss"
    whole.save(str(f), ids=np.arange(n))
    q = oracle.synth_queries(DIMS, 150, n, n_clusters=8)
    devs = _group_devices()
    from image_recommender_b200.sharded import shard_range
    grp = irb.ShardGroup(devs)
    shards = []
    for r, dev in enumerate(devs):
        r0, r1 = shard_range(n, len(devs), r)
        s = irb.FlatShard(DIMS, r1 - r0, device=dev, base_offset=r0)
        s.add_tables([t[r0:r1] for t in tabs])
        shards.append(s)
    grp.set_shards(shards)
    loaded = irb.ShardGroup.load(str(f), devs)
    assert grp.ntotal == n and loaded.ntotal == n and loaded.d == sum(DIMS)
    for nq, k in ((1, 10), (7, 5), (150, 10), (33, 40)):
        want = whole.search_ip(q[:nq], k)
        w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q[:nq], k, pk["norm2"])
        assert np.array_equal(want[1], w_lab)
        for g in (grp, loaded):
            got = g.search_ip(q[:nq], k)
            assert np.array_equal(got[1], want[1])
            assert np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32))
            assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32))
    offs = np.array([0, 1, 4, 6, 11], np.int32)
    a = whole.search_groups(q[:11], offs, 10)
    b = loaded.search_groups(q[:11], offs, 10)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    # resident-query form used by bench.py --single-process
    grp.put_queries(q[:64], 10)
    grp.run(64, 10)
    assert grp.last_run_ms() > 0
    got = grp.get_results(64, 10)
    want = whole.search_ip(q[:64], 10)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32))
    loaded.close()
    grp.close()
    whole.close()


def test_image_recommender_device_all(gpu, tmp_path, monkeypatch):
    """ImageRecommender(device='all'): the README flow on every visible GPU from one process, same hits as one GPU."""
    import sqlite3
    import pickle
    from main.create_index import FAISSIndexBuilderDB
    from main.search_from_image import ImageRecommender
    monkeypatch.chdir(tmp_path)
    (tmp_path / "image_data").mkdir()
    rng = np.random.default_rng(3)
    con = sqlite3.connect("images.db")
    con.executescript("CREATE TABLE images (id INTEGER PRIMARY KEY, path TEXT);"
                      "CREATE TABLE color_vectors (image_id INTEGER PRIMARY KEY, color_vector_blob BLOB);")
    for i in range(1, 301):
        v = np.abs(rng.normal(size=48)).astype(np.float32)
        con.execute("INSERT INTO images VALUES (?, ?)", (i, f"image_data/img_{i:04d}.jpg"))
        con.execute("INSERT INTO color_vectors VALUES (?, ?)", (i, sqlite3.Binary(pickle.dumps(v / np.linalg.norm(v), protocol=pickle.HIGHEST_PROTOCOL))))
    con.commit()
    con.close()
    FAISSIndexBuilderDB(db_path="images.db", vector_types=["color"], log_dir=str(tmp_path / "logs")).build_index()
    one = ImageRecommender(images_root="image_data", db_path="images.db", top_k=5, device=0)
    many = ImageRecommender(images_root="image_data", db_path="images.db", top_k=5, device="all")
    for i in (1, 57, 300):
        qp = [str(tmp_path / "image_data" / f"img_{i:04d}.jpg")]
        a, b = one.search_similar_images(qp, "color"), many.search_similar_images(qp, "color")
        assert a == b and a[0][0].name == f"img_{i:04d}.jpg"
    one.close()
    many.close()


@pytest.mark.parametrize("dims,n", [(DIMS, 20000), ([48], 40000), ([128], 30000)])
def test_transposed_kernel_matches_oracle(gpu, dims, n):
    """score_tn_kernel (DB rows on the MMA's M, queries on N; path 4) returns the oracle's bits for every batch
    size it serves (N = 16 ... 256 columns), with and without a seeded floor, through K-collect on a
    near-duplicate burst, and equals the M = queries kernels bit for bit."""
    import image_recommender_b200 as irb
    from image_recommender_b200 import _capi
    tabs, _ = _mk(n, dims)
    for t in tabs:
        t[5000:5200] = t[4000] * (1.0 + 1e-4 * np.arange(200, dtype=np.float32)[:, None])
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(dims, n, device=gpu)
    ix.add_tables(tabs)
    q = oracle.synth_queries(dims, 256, n, n_clusters=8)
    q[3] = pk["f32"][5100] / np.sqrt(np.float32(len(dims)))
    q = oracle.normalize_l2(q)
    for nq, k in ((1, 10), (5, 32), (16, 10), (17, 5), (100, 10), (130, 10), (160, 10), (200, 1), (255, 10), (256, 10)):
        w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q[:nq], k, pk["norm2"])
        ix.set_option(_capi.OPT_TN, 0)
        base = ix.search_ip(q[:nq], k)
        assert ix.stats()["path"] in (2, 3)
        for seed in (1, 0):
            ix.set_option(_capi.OPT_TN, 1)
            ix.set_option(_capi.OPT_SEED, seed)
            got = ix.search_ip(q[:nq], k)
            st = ix.stats()
            assert st["path"] == 4, st
            assert np.array_equal(got[1], w_lab), (nq, k, seed)
            assert np.array_equal(got[2].view(np.uint32), w_ip.view(np.uint32))
            assert np.array_equal(got[0].view(np.uint32), w_dist.view(np.uint32))
            assert np.array_equal(got[1], base[1]) and np.array_equal(got[2].view(np.uint32), base[2].view(np.uint32))
        ix.set_option(_capi.OPT_SEED, 1)
    ix.close()


def test_sampling_pass_one_wave_equals_main_grid_and_unseeded(gpu):
    """The seeding's sampling pass on its own grid (one wave of long CTAs over strided tiles, 1.5 sqrt(tiles) tiles)
    against the main pass's grid (every split its first tiles) and against no seeding at all: any subset of rows
    gives a valid floor, so ids and scores are bit-identical — on the single-CTA kernel (3 query tiles), the
    CTA-pair kernel (3 and 4 query tiles) and the transposed kernel (160 queries); a forced sample size too."""
    import torch
    import image_recommender_b200 as irb
    from image_recommender_b200 import _capi
    n, k = 2_000_000, 10
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.fill_synthetic(n, total_rows=n)
    q = ix.synth_queries_device(1000, total_rows=n)
    src = np.array([oracle.synth_query_source(0x5EED, i, n) for i in range(1000)])
    for m, path in ((160, 4), (300, 2), (700, 3), (1000, 3)):
        got = {}
        ix.set_option(_capi.OPT_TN, 1 if path == 4 else -1)   # (auto takes the transposed kernel from 2.5 M rows on)
        for name, wave, seed, launches in (("wave", 1, 1, 7), ("main_grid", 0, 1, 7), ("forced_sample", 1, 3, 7),
                                            ("unseeded", 1, 0, 5)):
            if path == 4 and seed == 0:
                continue                       # the transposed kernel only runs behind a seeded floor
            ix.set_option(_capi.OPT_SAMPLE_WAVE, wave)
            ix.set_option(_capi.OPT_SEED, seed)
            dist, lab, ip = ix.search_device(q[:m].contiguous(), k)
            torch.cuda.synchronize()
            st = ix.stats()
            assert st["launches"] == launches and st["path"] == path and st["n_uncertified"] == 0, (name, m, st)
            got[name] = (lab.cpu().numpy(), ip.cpu().numpy(), dist.cpu().numpy())
        assert np.array_equal(got["wave"][0][:, 0], src[:m])
        assert (np.diff(got["wave"][2], axis=1) >= 0).all()
        for name in got:
            assert np.array_equal(got["wave"][0], got[name][0]), (name, m)
            assert np.array_equal(got["wave"][1].view(np.uint32), got[name][1].view(np.uint32)), (name, m)
    ix.set_option(_capi.OPT_SAMPLE_WAVE, 1); ix.set_option(_capi.OPT_SEED, 1); ix.set_option(_capi.OPT_TN, -1)
    ix.close()


def test_group_rank_failure_is_reported_at_once(gpu):
    """A rank whose local search fails must not leave the root waiting out the exchange's 10 s peer timeout: it
    publishes its epoch without records (b2k_xchg_skip) and the group call returns that rank's own error; the
    next search, with the cause removed, is exact again."""
    import image_recommender_b200 as irb
    from image_recommender_b200 import _capi
    from image_recommender_b200.sharded import shard_range
    dims, n = [4160], 3000                                  # 4160 columns: beyond what K-scan supports
    tabs = oracle.synth_rows(dims, n, total_rows=n, n_clusters=8)
    pk = oracle.pack(tabs)
    q = oracle.synth_queries(dims, 5, n, n_clusters=8)
    devs = _group_devices()
    grp = irb.ShardGroup(devs)
    shards = []
    for r, dev in enumerate(devs):
        r0, r1 = shard_range(n, len(devs), r)
        s = irb.FlatShard(dims, r1 - r0, device=dev, base_offset=r0)
        s.add_tables([t[r0:r1] for t in tabs])
        shards.append(s)
    grp.set_shards(shards)
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
    assert np.array_equal(grp.search_ip(q, 10)[1], w_lab)
    shards[1].set_option(_capi.OPT_PATH, _capi.PATH_SCAN)   # rank 1 alone: "no scoring path supports Dp"
    t0 = time.perf_counter()
    with pytest.raises(_capi.B2KError) as ei:
        grp.search_ip(q, 10)
    assert time.perf_counter() - t0 < 3.0, "the root waited for the peer timeout"
    assert "rank 1" in str(ei.value) and "Dp" in str(ei.value), str(ei.value)
    shards[1].set_option(_capi.OPT_PATH, _capi.PATH_AUTO)
    got = grp.search_ip(q, 10)
    assert np.array_equal(got[1], w_lab) and np.array_equal(got[2].view(np.uint32), w_ip.view(np.uint32))
    grp.close()
