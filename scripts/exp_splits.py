"""GPU experiment: DB-split count / sample size of the CTA-pair scoring kernel vs shard size and batch."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi

DIMS = [48, 128, 1792]; D = sum(DIMS)
def run(rows, batches, variants, k=10):
    s = irb.FlatShard(DIMS, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    for B in batches:
        q = s.synth_queries_device(B, total_rows=rows)
        ref = None
        for sp, seed in variants:
            s.set_option(_capi.OPT_SPLITS, sp)
            s.set_option(_capi.OPT_SEED, seed)
            for _ in range(3): out = s.search_device(q, k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): out = s.search_device(q, k)
            e1.record(); torch.cuda.synchronize()
            st = s.stats()
            lab = out[1].clone()
            if ref is None: ref = lab
            same = bool(torch.equal(ref, lab))
            ms = e0.elapsed_time(e1) / 5
            print(json.dumps({"rows": rows, "B": B, "splits": sp, "seed": seed, "n_splits": st["n_splits"], "ms": round(ms, 3), "score_ms": round(st["score_ms"], 3),
                              "tail_ms": round(st["tail_ms"], 3), "tc_frac": round(2.0*B*rows*D/(st["score_ms"]*1e-3)/1e12/1642.5, 3),
                              "unc": st["n_uncertified"], "cand_per_q": round(st["n_candidates"]/max(st["n_queries"],1), 1),
                              "path": st["path"], "same": same}), flush=True)
    s.close()

V = [(0, 1), (148, 1), (148, 2), (74, 1), (74, 4), (74, 8), (37, 1), (37, 4), (37, 16)]
run(1_250_000, [4096, 512, 256], [(0, 1), (148, 1), (74, 1), (37, 1), (37, 2)])
run(10_000_000, [4096, 768, 512, 256], V)
