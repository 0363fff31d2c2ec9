"""GPU debug: fused tail vs staged tail on the heavy-collect scenario of test_more_queries_than_one_pass."""
import sys
sys.path.insert(0, ".")
import numpy as np
import oracle
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
DIMS = [48, 128, 1792]
n, k = 900, 7
tabs = oracle.synth_rows(DIMS, n, total_rows=n, n_clusters=8)
pk = oracle.pack(tabs)
ix = irb.FlatShard(DIMS, n, device=0)
ix.add_tables(tabs)
for nq in (300, 2000, 16384):
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=9)
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
    for fused in (0, 1, 1, 1):
        ix.set_option(_capi.OPT_FUSED_TAIL, fused)
        dist, lab, ip = ix.search_ip(q, k)
        st = ix.stats()
        bad = np.where((lab != w_lab).any(axis=1))[0]
        print(f"nq={nq} fused={fused} mismatching queries={len(bad)} first={bad[:8].tolist()} unc={st['n_uncertified']} sat={st['n_saturated']} cands={st['n_candidates']}", flush=True)
        if len(bad):
            b = bad[0]
            print("   got ", lab[b].tolist(), ip[b].tolist()); print("   want", w_lab[b].tolist(), w_ip[b].tolist(), flush=True)
