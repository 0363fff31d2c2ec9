"""GPU parity: libb2k.so (through the C ABI / FlatShard) against the CPU oracle, bit for bit.

Run with `pytest -m gpu` on a B200.  Sizes are chosen so the oracle finishes in seconds.
"""
import numpy as np
import pytest

import oracle
from conftest import DIMS

pytestmark = pytest.mark.gpu


def _irb():
    import image_recommender_b200 as irb
    return irb


def _mk(n, dims=DIMS, n_clusters=8, seed=0xC0FFEE):
    tabs = oracle.synth_rows(dims, n, total_rows=n, n_clusters=n_clusters, seed=seed)
    return tabs, oracle.pack(tabs)


@pytest.fixture(scope="module")
def db20k(gpu):
    irb = _irb()
    n = 20000
    tabs, pk = _mk(n)
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.add_tables(tabs)
    yield ix, pk, n
    ix.close()


# ------------------------------------------------------------------------------------ pack
@pytest.mark.parametrize("dims,n", [(DIMS, 3001), ([5, 3, 70], 257), ([48], 1000), ([1792], 129)])
def test_pack_bit_exact(gpu, dims, n):
    irb = _irb()
    tabs, pk = _mk(n, dims)
    ix = irb.FlatShard(dims, n, device=gpu)
    ix.add_tables(tabs[:])
    f, b, n2 = ix.get_rows(0, n)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32))
    assert np.array_equal(b, pk["bf16"])
    assert np.array_equal(n2.view(np.uint32), pk["norm2"].view(np.uint32))
    st = ix.stats()
    assert np.float32(st["err_max"]) == np.sqrt(pk["stats"][0])
    assert np.float32(st["norm_max"]) == np.sqrt(pk["stats"][1])
    ix.close()


def test_pack_chunked_add_and_zero_rows(gpu):
    """add() in ragged pieces (incl. an empty one and an all-zero row) == one add()."""
    irb = _irb()
    n = 1000
    tabs, _ = _mk(n)
    tabs[1][17] = 0.0                      # zero-norm part stays zero (no NaN)
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(DIMS, 16, device=gpu)   # grows through reserve()
    for lo, hi in [(0, 1), (1, 1), (1, 400), (400, 1000)]:
        ix.add_tables([t[lo:hi] for t in tabs])
    assert ix.ntotal == n
    f, b, n2 = ix.get_rows(0, n)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32))
    assert np.array_equal(b, pk["bf16"])
    assert np.isfinite(f).all()
    ix.close()


def test_add_concat_matches_tables(gpu):
    irb = _irb()
    tabs, pk = _mk(300)
    ix = irb.FlatShard(DIMS, 300, device=gpu)
    ix.add(np.concatenate(tabs, axis=1))      # the reference's index.add(arr) form
    f, _, _ = ix.get_rows(0, 300)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32))
    ix.close()


def test_normalize_l2(gpu):
    irb = _irb()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 1968)).astype(np.float32)
    x[5] = 0
    want = oracle.normalize_l2(x)
    got = x.copy()
    irb.normalize_L2(got, device=gpu)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_synthetic_generator_matches_oracle(gpu):
    irb = _irb()
    n, total = 5000, 77777
    ix = irb.FlatShard(DIMS, n, device=gpu, base_offset=1234)
    ix.fill_synthetic(n, total_rows=total, n_clusters=64)
    tabs = oracle.synth_rows(DIMS, n, first=1234, total_rows=total, n_clusters=64)
    pk = oracle.pack(tabs)
    f, b, _ = ix.get_rows(0, n)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32))
    assert np.array_equal(b, pk["bf16"])
    q = ix.synth_queries_device(33, total_rows=total, n_clusters=64).cpu().numpy()
    want = oracle.synth_queries(DIMS, 33, total, n_clusters=64)
    assert np.array_equal(q.view(np.uint32), want.view(np.uint32))
    ix.close()


# ------------------------------------------------------------------------------------ search
def _check(ix, pk, q, k, base=0):
    dist, lab, ip = ix.search_ip(q, k)
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"], base_offset=base)
    assert np.array_equal(lab, w_lab), (lab[:2], w_lab[:2])
    assert np.array_equal(ip.view(np.uint32), w_ip.view(np.uint32))
    assert np.array_equal(dist.view(np.uint32), w_dist.view(np.uint32))
    return ix.stats()


def _set_path(ix, name):
    """scan = K-scan (CUDA cores); tc = tcgen05 single-CTA tiles; tc2 = tcgen05 CTA pairs."""
    from image_recommender_b200 import _capi
    path, pair, want = {"scan": (1, -1, 1), "tc": (2, 0, 2), "tc2": (2, 1, 3), "auto": (0, -1, None)}[name]
    ix.set_option(_capi.OPT_PATH, path)
    ix.set_option(_capi.OPT_TC_PAIR, pair)
    return want


@pytest.mark.parametrize("path", ["scan", "tc", "tc2"])
@pytest.mark.parametrize("nq,k", [(1, 10), (3, 5), (4, 10), (7, 1), (130, 10), (300, 32), (513, 10)])
def test_search_matches_oracle(db20k, path, nq, k):
    from image_recommender_b200 import _capi
    ix, pk, n = db20k
    want = _set_path(ix, path)
    ix.set_option(_capi.OPT_FORCE_EXACT, 0)
    if path == "scan" and nq > 300:
        pytest.skip("K-scan serves tiny batches")
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=8)
    st = _check(ix, pk, q, k)
    assert st["path"] == want
    _set_path(ix, "auto")


@pytest.mark.parametrize("path", ["tc", "tc2"])
def test_seeded_threshold_matches_oracle_tc(db20k, path):
    """With 4 DB splits every split has ~20 tiles, so the sampling pass + seeded admission floor
    (DESIGN.md "seeding") is active; results must not change by a bit, with or without it."""
    from image_recommender_b200 import _capi
    ix, pk, n = db20k
    _set_path(ix, path)
    ix.set_option(_capi.OPT_SPLITS, 4)
    ix.set_option(_capi.OPT_INLINE_SEED, 0)            # the three-launch form (sampling pass, seed kernel, main pass)
    for nq, k in ((5, 10), (300, 10), (64, 32)):
        q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=1000 + nq)
        for seed in (1, 0):
            ix.set_option(_capi.OPT_SEED, seed)
            st = _check(ix, pk, q, k)
            assert st["launches"] == (7 if seed else 5)        # prep, [sample, seed,] score, tail, collect, exact
    ix.set_option(_capi.OPT_SEED, 1)
    ix.set_option(_capi.OPT_INLINE_SEED, 1)
    # in-kernel seeding where the grid is one wave (4 splits: 4 CTAs / 8 CTAs of 4 pairs): one scoring launch
    for nq in ((3, 4) if path == "tc" else (5, 16)):
        ix.set_option(_capi.OPT_SEED, 2)               # explicit sample size: three launches
        q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=2000 + nq)
        assert _check(ix, pk, q, 10)["launches"] == 7
        ix.set_option(_capi.OPT_SEED, 1)
        assert _check(ix, pk, q, 10)["launches"] == 5      # tc2: in-kernel seeding; tc: <= 4 queries run unseeded
    ix.set_option(_capi.OPT_SPLITS, 0)
    _set_path(ix, "auto")


def test_threshold_tightening_keeps_bits_and_cuts_candidates_tc(db20k):
    """The exact-score tightening of the candidate threshold (select_kernel) re-ranks fewer rows and
    returns the same bits."""
    from image_recommender_b200 import _capi
    ix, pk, n = db20k
    q = oracle.synth_queries(DIMS, 40, n, n_clusters=8, qseed=4242)
    cands = {}
    for path in ("scan", "tc2"):
        _set_path(ix, path)
        for on in (0, 1):
            ix.set_option(_capi.OPT_TIGHTEN, 2 * on)          # 2 = always (1 = auto: only from 32 queries)
            st = _check(ix, pk, q[:4] if path == "scan" else q, 10)
            cands[(path, on)] = st["n_candidates"] / st["n_queries"]
            assert st["n_uncertified"] == 0
    ix.set_option(_capi.OPT_TIGHTEN, 1)
    _set_path(ix, "auto")
    assert cands[("tc2", 1)] < 0.7 * cands[("tc2", 0)] and cands[("scan", 1)] < cands[("scan", 0)]
    assert cands[("tc2", 1)] >= 10


@pytest.mark.parametrize("path", ["scan", "tc", "tc2"])
@pytest.mark.parametrize("nq,k", [(1, 33), (5, 100), (130, 64), (300, 500), (3, 1024)])
def test_large_k_matches_oracle(db20k, path, nq, k):
    """k > 32 (faiss accepts any k): more DB splits, sorted per-query lists, paged exhaustive scan —
    same bits as the oracle on every path, certified or not."""
    from image_recommender_b200 import _capi
    ix, pk, n = db20k
    want = _set_path(ix, path)
    if path == "scan" and nq > 8:
        pytest.skip("K-scan serves tiny batches")
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=500 + k)
    st = _check(ix, pk, q, k)
    assert st["path"] == want
    if k in (33, 100):
        ix.set_option(_capi.OPT_FORCE_EXACT, 1)           # the paged exhaustive scan alone
        st = _check(ix, pk, q, k)
        assert st["n_uncertified"] == nq
        ix.set_option(_capi.OPT_FORCE_EXACT, 0)
    _set_path(ix, "auto")


def test_large_k_edge_cases(gpu):
    """k beyond the row count (padding), beyond the limit (error), ties across a page boundary of the
    exhaustive scan, and a cross-shard merge of k = 100 lists."""
    import torch
    irb = _irb()
    from image_recommender_b200 import _capi
    tabs, _ = _mk(700)
    for t in tabs:
        t[200:290] = t[5]                  # 91 identical rows: exact ties straddle results 32 / 64
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(DIMS, 700, device=gpu)
    ix.add_tables(tabs)
    q = oracle.normalize_l2(pk["f32"][[5, 250, 17]])
    for path in ("scan", "tc", "tc2"):
        _set_path(ix, path)
        for force in (0, 1):
            ix.set_option(_capi.OPT_FORCE_EXACT, force)
            _check(ix, pk, q, 100)
            dist, lab, ip = ix.search_ip(q, 1000)         # k > ntotal: -1 padding after 700 results
            assert (lab[:, 700:] == -1).all() and (lab[:, :700] >= 0).all()
            _check(ix, pk, q, 1000)
    ix.set_option(_capi.OPT_FORCE_EXACT, 0)
    with pytest.raises(irb.B2KError):
        ix.search(q, 1025)
    # two shards, k = 100, merged on the device
    a = irb.FlatShard(DIMS, 300, device=gpu, base_offset=0)
    b = irb.FlatShard(DIMS, 400, device=gpu, base_offset=300)
    a.add_tables([t[:300] for t in tabs]); b.add_tables([t[300:] for t in tabs])
    qd = torch.from_numpy(q).cuda(gpu)
    ra, rb = a.search_device(qd, 100), b.search_device(qd, 100)
    m_dist, m_lab, m_ip = irb.merge_topk_device(torch.stack([ra[2], rb[2]]).contiguous(), torch.stack([ra[0], rb[0]]).contiguous(),
                                                torch.stack([ra[1], rb[1]]).contiguous())
    torch.cuda.synchronize()
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, 100, pk["norm2"])
    assert np.array_equal(m_lab.cpu().numpy(), w_lab)
    assert np.array_equal(m_ip.cpu().numpy().view(np.uint32), w_ip.view(np.uint32))
    assert np.array_equal(m_dist.cpu().numpy().view(np.uint32), w_dist.view(np.uint32))
    ix.close(); a.close(); b.close()


def test_search_auto_path_and_fallback_tc(db20k):
    from image_recommender_b200 import _capi
    ix, pk, n = db20k
    _set_path(ix, "auto")
    q = oracle.synth_queries(DIMS, 200, n, n_clusters=8, qseed=99)
    assert _check(ix, pk, q[:1], 10)["path"] == 2        # tcgen05 at every batch size (measured faster)
    ix.set_option(_capi.OPT_SCAN_MAX_B, 4)
    assert _check(ix, pk, q[:3], 10)["path"] == 1        # K-scan on request
    ix.set_option(_capi.OPT_SCAN_MAX_B, 0)
    assert _check(ix, pk, q[:9], 10)["path"] == 2
    assert _check(ix, pk, q, 10)["path"] == 3            # 200 queries = one full 256-query pair tile
    q300 = oracle.synth_queries(DIMS, 300, n, n_clusters=8, qseed=98)
    assert _check(ix, pk, q300, 10)["path"] == 2         # 3 tiles of 128: a pair tile would be half empty
    q = q[:9]
    # exhaustive fp32 scan (the route of uncertified queries) returns the same bits
    ix.set_option(_capi.OPT_FORCE_EXACT, 1)
    st = _check(ix, pk, q, 10)
    assert st["n_uncertified"] == 9
    ix.set_option(_capi.OPT_FORCE_EXACT, 0)


def test_search_tight_candidate_budget_falls_back_tc(db20k):
    """With 32 candidate slots the budget overflows; results must still be exact."""
    from image_recommender_b200 import _capi
    ix, pk, n = db20k
    ix.set_option(_capi.OPT_RERANK, 32)
    q = oracle.synth_queries(DIMS, 6, n, n_clusters=8, qseed=7)
    for path in ("scan", "tc", "tc2"):
        _set_path(ix, path)
        st = _check(ix, pk, q, 10)
        assert st["n_uncertified"] > 0
    ix.set_option(_capi.OPT_RERANK, 0)
    _set_path(ix, "auto")


def test_self_query_rank0(db20k):
    """q = x/sqrt(T) (the cached-vector query, search_from_image.py:110-115): own offset first,
    ip = sqrt(T), L2^2 = (sqrt(T) - 1)^2."""
    ix, pk, n = db20k
    rows = np.array([0, 1, 4999, n - 1])
    q = oracle.normalize_l2(pk["f32"][rows])
    dist, lab, ip = ix.search_ip(q, 5)
    assert np.array_equal(lab[:, 0], rows)
    assert np.allclose(ip[:, 0], np.sqrt(3.0), atol=1e-5)
    assert np.allclose(dist[:, 0], (np.sqrt(3.0) - 1) ** 2, atol=1e-5)
    assert (np.diff(dist, axis=1) >= 0).all()       # ascending, as _fetch_results sorts


def test_k_larger_than_ntotal_pads_minus_one_tc(gpu):
    irb = _irb()
    tabs, pk = _mk(3)
    ix = irb.FlatShard(DIMS, 3, device=gpu)
    ix.add_tables(tabs)
    q = oracle.synth_queries(DIMS, 2, 3, n_clusters=8)
    for path in ("scan", "tc", "tc2"):
        _set_path(ix, path)
        dist, lab, ip = ix.search_ip(q, 5)
        assert (lab[:, 3:] == -1).all() and (lab[:, :3] >= 0).all()
        _check(ix, pk, q, 5)
    ix.close()


def test_empty_index_returns_padding(gpu):
    irb = _irb()
    ix = irb.FlatShard(DIMS, 8, device=gpu)
    q = oracle.synth_queries(DIMS, 2, 100, n_clusters=8)
    dist, lab = ix.search(q, 4)
    assert (lab == -1).all()
    ix.close()


def test_duplicate_rows_tie_break_by_offset_tc(gpu):
    """Exact-score ties (duplicate images) are ordered by lower offset, on every path.  41 identical
    rows saturate a partial list (32 entries): K-collect re-scans that DB split and the query stays
    certified; with K-collect off the same bits come from the exhaustive scan."""
    irb = _irb()
    from image_recommender_b200 import _capi
    tabs, _ = _mk(600)
    for t in tabs:
        t[100:141] = t[7]                 # 42 identical rows: more than one partial list holds
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(DIMS, 600, device=gpu)
    ix.add_tables(tabs)
    q = oracle.normalize_l2(pk["f32"][[7, 120]])
    for path in ("scan", "tc", "tc2"):
        _set_path(ix, path)
        ix.set_option(_capi.OPT_SPLITS, 0 if path == "scan" else 1)     # one split: its list must saturate
        for collect in (1, 0):
            ix.set_option(_capi.OPT_COLLECT, collect)
            st = _check(ix, pk, q, 32)
            if path != "scan":
                assert st["n_saturated"] == (2 if collect else 0)
                assert st["n_uncertified"] == (0 if collect else 2)
    ix.close()


@pytest.mark.parametrize("path", ["scan", "tc", "tc2"])
def test_near_duplicate_runs_are_collected(gpu, path):
    """A burst of 300 near-identical images stored next to each other (one DB split) and a second
    burst of 5000 (more than the candidate buffer holds): the first is served by K-collect, the
    second overflows into the exhaustive scan; both return the oracle's bits."""
    irb = _irb()
    from image_recommender_b200 import _capi
    n = 9000
    tabs, _ = _mk(n)
    rng = np.random.default_rng(5)
    for t in tabs:
        t[1000:1300] = t[1000] + 1e-4 * rng.standard_normal((300, t.shape[1])).astype(np.float32)
        t[3000:8000] = t[3000]
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.add_tables(tabs)
    _set_path(ix, path)
    q = oracle.normalize_l2(pk["f32"][[1100, 17, 3500]])
    st = _check(ix, pk, q, 10)
    assert st["n_saturated"] >= 2            # both bursts saturate at least one list
    assert st["n_uncertified"] == 1          # only the 5000-row burst needs the exhaustive scan
    q1 = oracle.normalize_l2(pk["f32"][[1100, 17]])
    st = _check(ix, pk, q1, 10)
    assert st["n_saturated"] >= 1 and st["n_uncertified"] == 0
    ix.close()


@pytest.mark.parametrize("dims,n", [([32768], 700), ([4096, 64], 1500), ([128], 5000), ([5, 3, 70], 3000), ([1], 500)])
def test_other_baseline_dims_match_oracle(gpu, dims, n):
    """BASELINE configs 4 / 4s: the raw 32768-d SIFT-VLAD descriptor, a wide two-table combo, and the
    stored 128-d table.  Pack, batch-1 and batch-130 search bit-equal to the oracle."""
    irb = _irb()
    tabs, pk = _mk(n, dims)
    ix = irb.FlatShard(dims, n, device=gpu)
    ix.add_tables(tabs)
    f, b, n2 = ix.get_rows(0, n)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32)) and np.array_equal(b, pk["bf16"])
    q = oracle.synth_queries(dims, 130, n, n_clusters=8, qseed=77)
    st1 = _check(ix, pk, q[:1], 10)
    assert st1["path"] == 2
    st = _check(ix, pk, q, 10)
    assert st["path"] == 3 and st["n_uncertified"] == 0
    ix.close()


def test_base_offset_and_two_shard_merge(gpu):
    import torch
    irb = _irb()
    n, k, nq = 6000, 10, 70
    tabs, pk = _mk(n)
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=3)
    cut = 2500
    a = irb.FlatShard(DIMS, cut, device=gpu, base_offset=0)
    b = irb.FlatShard(DIMS, n - cut, device=gpu, base_offset=cut)
    a.add_tables([t[:cut] for t in tabs])
    b.add_tables([t[cut:] for t in tabs])
    from image_recommender_b200 import _capi
    a.set_option(_capi.OPT_PATH, 1); b.set_option(_capi.OPT_PATH, 1)
    qd = torch.from_numpy(q).cuda(gpu)
    ra = a.search_device(qd, k)
    rb = b.search_device(qd, k)
    torch.cuda.synchronize()
    ip = torch.stack([ra[2], rb[2]]).contiguous()
    dist = torch.stack([ra[0], rb[0]]).contiguous()
    lab = torch.stack([ra[1], rb[1]]).contiguous()
    m_dist, m_lab, m_ip = irb.merge_topk_device(ip, dist, lab)
    torch.cuda.synchronize()
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
    assert np.array_equal(m_lab.cpu().numpy(), w_lab)
    assert np.array_equal(m_ip.cpu().numpy().view(np.uint32), w_ip.view(np.uint32))
    assert np.array_equal(m_dist.cpu().numpy().view(np.uint32), w_dist.view(np.uint32))
    # and the oracle's own merge agrees with it
    o_dist, o_lab, o_ip = oracle.merge_topk(ip.cpu().numpy(), dist.cpu().numpy(), lab.cpu().numpy())
    assert np.array_equal(o_lab, w_lab)
    a.close(); b.close()


def test_save_load_round_trip(gpu, tmp_path, db20k):
    irb = _irb()
    ix, pk, n = db20k
    ids = np.arange(n, dtype=np.int64) * 3 + 11
    p = tmp_path / "index_hnsw_color_sift_dreamsim.faiss"
    ix.save(p, ids)
    info = irb.file_info(p)
    assert info == {"n_rows": n, "table_dims": DIMS, "has_ids": True}
    assert np.array_equal(irb.load_ids(p, 5, 10), ids[5:15])
    whole = irb.FlatShard.load(p, device=gpu)
    part = irb.FlatShard.load(p, device=gpu, row_begin=5000, row_end=12000)
    assert whole.ntotal == n and part.ntotal == 7000 and part.base_offset == 5000
    f, b, n2 = whole.get_rows(0, n)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32))
    assert np.array_equal(b, pk["bf16"])
    assert np.array_equal(n2.view(np.uint32), pk["norm2"].view(np.uint32))
    q = oracle.synth_queries(DIMS, 4, n, n_clusters=8, qseed=11)
    _check(whole, pk, q, 10)
    sub = {k_: v[5000:12000] for k_, v in pk.items() if k_ != "stats"}
    _check(part, sub, q, 10, base=5000)
    whole.close(); part.close()


# ------------------------------------------------------------------- BASELINE-scale properties
def test_large_scale_properties_tc(gpu):
    """2M combo rows generated on the device (too large for the oracle): every query is a noisy
    copy of a seeded DB row, so rank 0 must be that row on every path; distances ascend; a
    two-shard split merged equals the single shard."""
    import torch
    irb = _irb()
    from image_recommender_b200 import _capi
    n, nq, k = 2_000_000, 256, 10
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.fill_synthetic(n, total_rows=n)
    q = ix.synth_queries_device(nq, total_rows=n)
    src = np.array([oracle.synth_query_source(0x5EED, i, n) for i in range(nq)])
    res = {}
    for path, m in (("scan", 8), ("tc", nq), ("tc2", nq)):
        _set_path(ix, path)
        dist, lab, ip = ix.search_device(q[:m].contiguous(), k)
        torch.cuda.synchronize()
        lab = lab.cpu().numpy(); dist = dist.cpu().numpy()
        assert np.array_equal(lab[:, 0], src[:m])
        assert (np.diff(dist, axis=1) >= 0).all()
        assert ix.stats()["n_uncertified"] <= m // 10
        res[path] = (lab, ip.cpu().numpy())
    for other in ("tc", "tc2"):
        assert np.array_equal(res["scan"][0], res[other][0][:8])
        assert np.array_equal(res["scan"][1].view(np.uint32), res[other][1][:8].view(np.uint32))
    assert np.array_equal(res["tc"][0], res["tc2"][0])
    # in-kernel seeding (single-CTA kernel, <= 128 queries: the CTAs exchange their lists after the first
    # tile and continue with a global floor) against the three-launch seeding and no seeding at all
    from image_recommender_b200 import _capi
    _set_path(ix, "tc")
    got = {}
    for name, inline, seed, launches in (("inline", 1, 1, 5), ("three_launch", 0, 1, 7), ("unseeded", 0, 0, 5)):
        ix.set_option(_capi.OPT_INLINE_SEED, inline)
        ix.set_option(_capi.OPT_SEED, seed)
        for m in (5, 64, 128):
            dist, lab, ip = ix.search_device(q[:m].contiguous(), k)
            torch.cuda.synchronize()
            st = ix.stats()
            assert st["launches"] == launches and st["n_uncertified"] == 0
            got[(name, m)] = (lab.cpu().numpy(), ip.cpu().numpy(), st["n_candidates"])
    for m in (5, 64, 128):
        for name in ("three_launch", "unseeded"):
            assert np.array_equal(got[("inline", m)][0], got[(name, m)][0])
            assert np.array_equal(got[("inline", m)][1].view(np.uint32), got[(name, m)][1].view(np.uint32))
        assert np.array_equal(got[("inline", m)][0], res["tc2"][0][:m])
    ix.set_option(_capi.OPT_INLINE_SEED, 1); ix.set_option(_capi.OPT_SEED, 1)
    # the CTA-pair kernel does the same for one query tile (<= 256 queries) when its pairs fill one wave
    _set_path(ix, "tc2")
    for m in (130, 256):
        outs = []
        for inline, launches in ((1, 5), (0, 7)):
            ix.set_option(_capi.OPT_INLINE_SEED, inline)
            dist, lab, ip = ix.search_device(q[:m].contiguous(), k)
            torch.cuda.synchronize()
            st = ix.stats()
            assert st["launches"] == launches and st["path"] == 3 and st["n_uncertified"] == 0
            outs.append((lab.cpu().numpy(), ip.cpu().numpy()))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
        assert np.array_equal(outs[0][0], res["tc2"][0][:m])
    ix.set_option(_capi.OPT_INLINE_SEED, 1)
    ix.close()


def test_faiss_shim_protocol(gpu, tmp_path):
    """The members of the faiss namespace the reference touches, served by the engine
    (INTEGRATION.md §2): construction, hnsw params, is_trained/train, add, ntotal, search,
    write_index/read_index, normalize_L2."""
    import image_recommender_b200.faiss_shim as faiss
    tabs, pk = _mk(700)
    arr = np.concatenate(tabs, axis=1)
    index = faiss.IndexIVFPQ(faiss.IndexHNSWFlat(1968, 32), 1968, 2048, 48, 12)
    index.table_dims = DIMS
    index.hnsw.efConstruction = 200
    assert index.is_trained and index.ntotal == 0
    index.train(arr)
    index.add(arr[:300]); index.add(arr[300:])
    assert index.ntotal == 700
    q = oracle.synth_queries(DIMS, 1, 700, n_clusters=8)
    qq = q.copy(); faiss.normalize_L2(qq)
    d, i = index.search(qq, 5)
    wd, wi, _ = oracle.search_exact(pk["f32"], oracle.normalize_l2(q), 5, pk["norm2"])
    assert np.array_equal(i, wi) and np.array_equal(d.view(np.uint32), wd.view(np.uint32))
    faiss.write_index(index, str(tmp_path / "x.faiss"))
    back = faiss.read_index(str(tmp_path / "x.faiss"))
    d2, i2 = back.search(qq, 5)
    assert back.ntotal == 700 and np.array_equal(i2, wi) and np.array_equal(d2.view(np.uint32), wd.view(np.uint32))


# ------------------------------------------------------------------- the other BASELINE configs
@pytest.mark.parametrize("dims,n,nq,k", [
    ([48], 10000, 1, 5),          # config 1: color-only, top-5 (the reference's own __main__ case)
    ([1792], 6000, 200, 10),      # config 2: DreamSim-only table, query batch
    ([32768], 1500, 1, 10),       # config 4: raw SIFT-VLAD (32768-d): beyond K-scan's row width -> tcgen05 path
    ([32768], 1500, 3, 10),
    ([128, 1792], 3000, 9, 10),   # a two-table combo in a non-canonical order
])
def test_other_baseline_configs(gpu, dims, n, nq, k):
    irb = _irb()
    tabs = oracle.synth_rows(dims, n, total_rows=n, n_clusters=8, abs_mask=1 if dims[0] == 48 else 0)
    pk = oracle.pack(tabs)
    ix = irb.FlatShard(dims, n, device=gpu)
    ix.add_tables(tabs)
    f, b, _ = ix.get_rows(0, n)
    assert np.array_equal(f.view(np.uint32), pk["f32"].view(np.uint32)) and np.array_equal(b, pk["bf16"])
    q = oracle.synth_queries(dims, nq, n, n_clusters=8, abs_mask=1 if dims[0] == 48 else 0)
    st = _check(ix, pk, q, k)
    assert st["path"] == (3 if nq > 128 else 2)
    ix.close()


def test_peer_exchange_two_ranks_one_process(gpu):
    """K-exchange (csrc/xchg.cu) with both 'ranks' on one GPU in one process (plain pointers instead
    of IPC handles).  Push and merge are separate launches, so pushing both ranks before merging
    never makes a kernel wait for a later launch.  Three epochs exercise the double buffering."""
    import torch
    irb = _irb()
    from image_recommender_b200 import _capi
    from image_recommender_b200.sharded import PeerExchange
    n, k, cut = 5000, 10, 2100
    tabs, pk = _mk(n)
    a = irb.FlatShard(DIMS, cut, device=gpu, base_offset=0)
    b = irb.FlatShard(DIMS, n - cut, device=gpu, base_offset=cut)
    a.add_tables([t[:cut] for t in tabs]); b.add_tables([t[cut:] for t in tabs])
    xs = [PeerExchange(gpu, r, 2, max_entries=64 * 32, _local_peers=True) for r in range(2)]
    for x in xs:
        x.connect_local([y.raw_ptr for y in xs])
    for epoch, nq in enumerate((1, 37, 64)):
        q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=50 + epoch)
        qd = torch.from_numpy(q).cuda(gpu)
        ra = a.search_device(qd, k)
        rb = b.search_device(qd, k)
        xs[0].push(ra[2], ra[0], ra[1])
        xs[1].push(rb[2], rb[0], rb[1])
        outs = [x.merge(nq, k) for x in xs]
        torch.cuda.synchronize()
        w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
        for o_d, o_l, o_ip in outs:
            assert np.array_equal(o_l.cpu().numpy(), w_lab)
            assert np.array_equal(o_ip.cpu().numpy().view(np.uint32), w_ip.view(np.uint32))
            assert np.array_equal(o_d.cpu().numpy().view(np.uint32), w_dist.view(np.uint32))
    for x in xs:
        x.close()
    a.close(); b.close()


def test_config5_per_gpu_capacity_slice(gpu):
    """BASELINE config 5 is 100 M DreamSim rows over 8 GPUs = 12.5 M x 1792 per GPU: 44.8 GB bf16 +
    89.6 GB fp32 resident on ONE GPU (HBM-capacity sizing, 64-bit byte offsets).  Queries are noisy
    copies of seeded rows spread over the whole shard, so rank 0 must be the source row."""
    import torch
    irb = _irb()
    free, _total = torch.cuda.mem_get_info(gpu)
    n = 12_500_000
    need = n * 1792 * 6 + (4 << 30)
    if free < need:
        pytest.skip(f"needs {need >> 30} GiB free HBM, {free >> 30} GiB available")
    ix = irb.FlatShard([1792], n, device=gpu, base_offset=87_500_000)      # the last of 8 shards
    ix.fill_synthetic(n, total_rows=100_000_000, abs_mask=0)
    assert ix.ntotal == n
    # queries drawn from THIS shard's rows: sources in [base, base + n)
    q = ix.synth_queries_device(512, total_rows=100_000_000, abs_mask=0)
    src = np.array([oracle.synth_query_source(0x5EED, i, 100_000_000) for i in range(512)])
    mine = (src >= 87_500_000) & (src < 100_000_000)
    assert mine.sum() > 20
    for m in (1, 512):
        dist, lab, ip = ix.search_device(q[:m].contiguous(), 10)
        torch.cuda.synchronize()
        lab = lab.cpu().numpy()
        assert (lab[:, 0][mine[:m]] == src[:m][mine[:m]]).all()
        assert (lab >= 87_500_000).all() and (lab < 100_000_000).all()
    # the last row of the shard is addressable (64-bit offsets) and finds itself
    f, _, _ = ix.get_rows(n - 1, 1)
    d, l = ix.search(f, 1)
    assert l[0, 0] == 99_999_999 and d[0, 0] < 1e-5
    ix.close()


def test_sharded_index_single_rank_and_row_ranges(gpu, tmp_path, db20k):
    """ShardedIndex (the faiss-protocol object ImageRecommender uses under torchrun): with one rank it is
    the whole file; two row-range loads of the same file merged by K-merge give the same bits — the
    arithmetic of the N > 1 deployment without needing N GPUs (the real thing: scripts/check_sharded.py)."""
    import torch
    irb = _irb()
    from image_recommender_b200.sharded import ShardedIndex, shard_range
    ix, pk, n = db20k
    p = tmp_path / "index_hnsw_color_sift_dreamsim.faiss"
    ix.save(p, np.arange(n, dtype=np.int64))
    q = oracle.synth_queries(DIMS, 37, n, n_clusters=8, qseed=321)
    want = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
    sh = ShardedIndex.load(p, device=gpu)
    assert sh.ntotal == n and sh.d == sum(DIMS)
    d, l = sh.search(q, 10)
    assert np.array_equal(l, want[1]) and np.array_equal(d.view(np.uint32), want[0].view(np.uint32))
    sh.close()
    parts = []
    qd = torch.from_numpy(q).cuda(gpu)
    for r in range(3):
        r0, r1 = shard_range(n, 3, r)
        part = irb.FlatShard.load(p, device=gpu, row_begin=r0, row_end=r1)
        assert part.base_offset == r0 and part.ntotal == r1 - r0
        parts.append([t.clone() for t in part.search_device(qd, 10)])
        part.close()
    m_dist, m_lab, m_ip = irb.merge_topk_device(torch.stack([x[2] for x in parts]).contiguous(),
                                                torch.stack([x[0] for x in parts]).contiguous(),
                                                torch.stack([x[1] for x in parts]).contiguous())
    torch.cuda.synchronize()
    assert np.array_equal(m_lab.cpu().numpy(), want[1])
    assert np.array_equal(m_dist.cpu().numpy().view(np.uint32), want[0].view(np.uint32))
    assert np.array_equal(m_ip.cpu().numpy().view(np.uint32), want[2].view(np.uint32))


def test_more_queries_than_one_pass(gpu):
    """nq beyond the per-pass workspace limit (16384): the host and device entry points split the batch;
    results equal those of separate calls (and the oracle on a sample)."""
    import torch
    irb = _irb()
    n, k = 900, 7
    tabs, pk = _mk(n)
    ix = irb.FlatShard(DIMS, n, device=gpu)
    ix.add_tables(tabs)
    nq = 16384 + 300
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=9)
    dist, lab, ip = ix.search_ip(q, k)
    d1, l1, i1 = ix.search_ip(q[:16384], k)
    d2, l2, i2 = ix.search_ip(q[16384:], k)
    assert np.array_equal(lab, np.concatenate([l1, l2])) and np.array_equal(dist.view(np.uint32), np.concatenate([d1, d2]).view(np.uint32))
    sample = np.r_[0:50, 16380:16390, nq - 40:nq]
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q[sample], k, pk["norm2"])
    assert np.array_equal(lab[sample], w_lab) and np.array_equal(ip[sample].view(np.uint32), w_ip.view(np.uint32))
    qd = torch.from_numpy(q).cuda(gpu)
    dd, ld, _ = ix.search_device(qd, k)
    torch.cuda.synchronize()
    assert np.array_equal(ld.cpu().numpy(), lab) and np.array_equal(dd.cpu().numpy().view(np.uint32), dist.view(np.uint32))
    ix.close()


def test_save_shard_equals_whole_save(gpu, tmp_path):
    """b2k_save_shard: three shards written into one laid-out file (creator first) == b2k_save of the whole."""
    irb = _irb()
    from image_recommender_b200.sharded import shard_range
    n = 1001
    tabs, _ = _mk(n)
    ids = np.arange(n, dtype=np.int64) * 7 + 3
    whole = irb.FlatShard(DIMS, n, device=gpu)
    whole.add_tables(tabs)
    whole.save(tmp_path / "whole.faiss", ids)
    for with_ids in (True, False):
        out = tmp_path / f"sharded_{with_ids}.faiss"
        for r in range(3):
            r0, r1 = shard_range(n, 3, r)
            part = irb.FlatShard(DIMS, r1 - r0, device=gpu, base_offset=r0)
            part.add_tables([t[r0:r1] for t in tabs])
            part.save_shard(out, ids[r0:r1] if with_ids else None, r0, n, create=(r == 0))
            part.close()
        if with_ids:
            assert out.read_bytes() == (tmp_path / "whole.faiss").read_bytes()
        else:
            whole.save(tmp_path / "whole_noids.faiss")
            assert out.read_bytes() == (tmp_path / "whole_noids.faiss").read_bytes()
    with pytest.raises(irb.B2KError):        # a shard that does not fit the layout
        whole.save_shard(tmp_path / "sharded_True.faiss", ids, 5, n, create=False)
    whole.close()
