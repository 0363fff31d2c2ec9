"""GPU debug: test_more_queries_than_one_pass, verbose."""
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
for fused in (1, 0):
    ix = irb.FlatShard(DIMS, n, device=0)
    ix.add_tables(tabs)
    ix.set_option(_capi.OPT_FUSED_TAIL, fused)
    nq = 16384 + 300
    q = oracle.synth_queries(DIMS, nq, n, n_clusters=8, qseed=9)
    w_dist, w_lab, w_ip = oracle.search_exact(pk["f32"], q, k, pk["norm2"])
    for rep in range(3):
        dist, lab, ip = ix.search_ip(q, k)
        st = ix.stats()
        bad = np.where((lab != w_lab).any(axis=1) | (dist.view(np.uint32) != w_dist.view(np.uint32)).any(axis=1))[0]
        print(f"fused={fused} rep={rep} combined call: bad={len(bad)} first={bad[:10].tolist()} last={bad[-3:].tolist()} unc={st['n_uncertified']} sat={st['n_saturated']}", flush=True)
        if len(bad):
            b = bad[0]
            print("   got ", lab[b].tolist(), dist[b].tolist()); print("   want", w_lab[b].tolist(), w_dist[b].tolist(), flush=True)
    d1, l1, i1 = ix.search_ip(q[:16384], k)
    d2, l2, i2 = ix.search_ip(q[16384:], k)
    print("separate: bad1", int((l1 != w_lab[:16384]).any(axis=1).sum()), "bad2", int((l2 != w_lab[16384:]).any(axis=1).sum()), flush=True)
    ix.close()
