"""The oracle against independent numpy restatements and its own invariants (CPU only)."""
import numpy as np
import pytest

import oracle
from conftest import DIMS


def test_bf16_rne_matches_torch():
    import torch
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(5000).astype(np.float32) * s for s in (1e-3, 1, 1e3)])
    x = np.concatenate([x, np.array([0.0, -0.0, np.inf, -np.inf, 1.00390625, 1.01171875], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(oracle.bf16_rne(x), want)
    assert np.array_equal(oracle.bf16_to_f32(want), torch.from_numpy(x).to(torch.bfloat16).float().numpy())


def test_sumsq_and_normalize_close_to_numpy():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((64, 1968)).astype(np.float32)
    x[3] = 0
    y = oracle.normalize_l2(x)
    ref = x / np.maximum(np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True), 1e-300)
    ref[3] = 0
    assert np.allclose(y, ref, rtol=0, atol=2e-7)
    assert (y[3] == 0).all()
    assert abs(float(oracle.sumsq32(x[0])) - float((x[0].astype(np.float64) ** 2).sum())) < 1e-3


def test_pack_semantics():
    tabs = oracle.synth_rows(DIMS, 500, n_clusters=4)
    pk = oracle.pack(tabs)
    f = pk["f32"]
    off = 0
    for d in DIMS:                                     # every part unit-norm (F2), row norm sqrt(T)
        assert np.allclose(np.linalg.norm(f[:, off:off + d].astype(np.float64), axis=1), 1.0, atol=1e-6)
        off += d
    assert np.allclose(pk["norm2"], 3.0, atol=1e-5)
    assert (f[:, :48] >= 0).all()                      # colour histograms are non-negative
    assert np.array_equal(pk["bf16"][:, :1968], oracle.bf16_rne(f))
    assert (pk["bf16"][:, 1968:] == 0).all() and pk["bf16"].shape[1] == 1984
    # no-normalise mode is a pure concat (the reference's _process_batch, create_index.py:160-189)
    raw = oracle.pack(tabs, normalize=False)["f32"]
    assert np.array_equal(raw, np.concatenate(tabs, axis=1))
    # normalising already-normalised parts is idempotent up to a few ulp
    parts = [f[:, :48], f[:, 48:176], f[:, 176:]]
    again = oracle.pack(parts)["f32"]
    assert np.abs(again.view(np.int32).astype(np.int64) - f.view(np.int32).astype(np.int64)).max() <= 4


def test_exact_search_vs_numpy_f64():
    n, nq, k = 3000, 17, 10
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=4)
    dist, lab, ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
    s = q.astype(np.float64) @ pk["f32"].astype(np.float64).T
    for i in range(nq):
        order = np.lexsort((np.arange(n), -s[i].astype(np.float32)))[:k]
        assert np.array_equal(lab[i], order)
        assert np.allclose(ip[i], s[i][order], rtol=0, atol=1e-7)
    # squared L2 of unit query vs sqrt(3)-norm rows: 1 + 3 - 2 ip, ascending
    assert np.allclose(dist, 4.0 - 2.0 * ip, atol=2e-6)
    assert (np.diff(dist, axis=1) >= 0).all()
    src = [oracle.synth_query_source(0x5EED, i, n) for i in range(nq)]
    assert np.array_equal(lab[:, 0], src)


def test_k_greater_than_n_and_merge():
    pk = oracle.pack(oracle.synth_rows(DIMS, 7, n_clusters=2))
    q = oracle.synth_queries(DIMS, 3, 7, n_clusters=2)
    dist, lab, ip = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
    assert (lab[:, 7:] == -1).all() and (lab[:, :7] >= 0).all()
    # shard-count invariance of the merged result (SURVEY §8c property)
    n = 1000
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
    q = oracle.synth_queries(DIMS, 9, n, n_clusters=4)
    want = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
    for G in (2, 4, 8):
        per = -(-n // G)
        parts = [oracle.search_exact(pk["f32"][g * per:(g + 1) * per], q, 10, pk["norm2"][g * per:(g + 1) * per],
                                     base_offset=g * per) for g in range(G)]
        m = oracle.merge_topk(np.stack([p[2] for p in parts]), np.stack([p[0] for p in parts]),
                              np.stack([p[1] for p in parts]))
        for a, b in zip(m, want):
            assert np.array_equal(a, b)


def test_certificate_bound_holds_on_cpu():
    """The bound the CUDA query-prep kernel uses (DESIGN.md 'certificate') really dominates the
    bf16 scoring error on this data (fp64 evaluation of the rounded operands)."""
    n = 4000
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
    q = oracle.synth_queries(DIMS, 8, n, n_clusters=4)
    x = pk["f32"].astype(np.float64)
    xb = oracle.bf16_to_f32(pk["bf16"])[:, :1968].astype(np.float64)
    qb = oracle.bf16_to_f32(oracle.bf16_rne(q)).astype(np.float64)
    E, X = np.sqrt(pk["stats"][0]), np.sqrt(pk["stats"][1])
    for i in range(8):
        err_tc = np.abs(xb @ qb[i] - x @ q[i].astype(np.float64)).max()
        err_scan = np.abs(xb @ q[i].astype(np.float64) - x @ q[i].astype(np.float64)).max()
        eps_tc = np.linalg.norm(qb[i]) * E + np.linalg.norm(qb[i] - q[i]) * (X + E)
        eps_scan = np.linalg.norm(q[i]) * E
        assert err_tc <= eps_tc and err_scan <= eps_scan
        assert eps_tc < 0.02          # and it is tight enough to be useful (cluster spread ~0.01)


def test_cpu_flat_port_agrees_with_oracle_ids():
    from oracle import cpu_flat
    n = 5000
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
    for nq in (3, 40):
        q = oracle.synth_queries(DIMS, nq, n, n_clusters=4)
        s, i = cpu_flat.search_flat_ip(pk["f32"], q, 10)
        _, lab, ip = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
        # identical up to fp32 near-ties of the BLAS summation order
        same = (i == lab).mean()
        assert same > 0.97
        assert np.allclose(s, ip, atol=2e-6)


def test_hnsw_restatement_recall():
    """The reference's IndexHNSWFlat(M=32, efC=200) restated (oracle/hnsw_ref.c): squared-L2
    results, ascending, near-perfect recall on an easy clustered set, monotone in efSearch."""
    from oracle import hnsw_ref
    n = 3000
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=16))
    q = oracle.synth_queries(DIMS, 50, n, n_clusters=16)
    dist, lab, _ = oracle.search_exact(pk["f32"], q, 10, pk["norm2"])
    ix = hnsw_ref.IndexHNSWFlat(1968, 32, 200, 64)
    ix.add(pk["f32"])
    rec = {}
    for ef in (10, 64):
        ix.efSearch = ef
        d, l = ix.search(q, 10)
        rec[ef] = np.mean([len(set(l[i]) & set(lab[i])) / 10 for i in range(50)])
        assert (np.diff(d, axis=1) >= 0).all()
    assert rec[64] >= 0.95 and rec[64] >= rec[10] - 1e-9
    same = l == lab
    assert np.allclose(d[same], dist[same], atol=1e-4)      # METRIC_L2 values, as index.search returns


def test_oracle_agrees_with_third_party_exact_knn():
    """Independent implementations of the same exact search (scikit-learn brute-force kNN on the
    squared-L2 metric the reference's index uses, scipy cdist in fp64): the oracle's ids are theirs
    wherever the fp64 gap between consecutive neighbours exceeds fp32 resolution."""
    from scipy.spatial.distance import cdist
    from sklearn.neighbors import NearestNeighbors
    n, nq, k = 4000, 25, 10
    pk = oracle.pack(oracle.synth_rows(DIMS, n, n_clusters=4))
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=4)
    dist, lab, ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(pk["f32"].astype(np.float64))
    d_sk, i_sk = nn.kneighbors(q.astype(np.float64))
    d64 = cdist(q.astype(np.float64), pk["f32"].astype(np.float64), "sqeuclidean")
    checked = 0
    for r in range(nq):
        order = np.argsort(d64[r], kind="stable")[:k + 1]
        gaps = np.diff(d64[r][order])
        clear = gaps > 1e-5                      # neighbours separated by more than fp32 rounding
        for j in range(k):
            if clear[j] and (j == 0 or clear[j - 1]):
                assert lab[r, j] == order[j] == i_sk[r, j]
                checked += 1
        assert np.allclose(dist[r], d64[r][lab[r]], atol=2e-5)
        assert np.allclose(d_sk[r], d64[r][i_sk[r]], atol=1e-9)
    assert checked > 0.8 * nq * k
