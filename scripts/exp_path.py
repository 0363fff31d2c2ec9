"""GPU experiment: K-scan vs K-score (seeding on/off) at tiny batches on the combo workload, per shard size."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
DIMS = [48, 128, 1792]; D = sum(DIMS)
for rows in (1_250_000, 10_000_000):
    s = irb.FlatShard(DIMS, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    for B in (1, 4, 8, 32, 128):
        q = s.synth_queries_device(B, total_rows=rows)
        ref = None
        for path, seed, inline in ((1, 1, 0), (2, 1, 1), (2, 1, 0), (2, 0, 0)):
            if path == 1 and B > 4: continue
            s.set_option(_capi.OPT_PATH, path); s.set_option(_capi.OPT_SEED, seed); s.set_option(_capi.OPT_INLINE_SEED, inline)
            for _ in range(5): out = s.search_device(q, 10)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): out = s.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            st = s.stats(); lab = out[1].clone()
            if ref is None: ref = lab
            print(json.dumps({"rows": rows, "B": B, "path": st["path"], "seed": seed, "inline": inline, "launches": st["launches"], "ms": round(e0.elapsed_time(e1) / 20, 4),
                              "score_ms": round(st["score_ms"], 4), "tail_ms": round(st["tail_ms"], 4),
                              "hbm_frac_total": round(2.0 * rows * D / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9 / 6558.1, 3),
                              "unc": st["n_uncertified"], "same": bool(torch.equal(ref, lab))}), flush=True)
    s.close()
