"""GPU experiment: in-kernel seeding in the CTA-pair kernel, 129..256 queries on an 8-GPU-sized shard."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
DIMS = [48, 128, 1792]
for rows in (1_250_000, 2_500_000):
    s = irb.FlatShard(DIMS, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    for B in (160, 256):
        q = s.synth_queries_device(B, total_rows=rows)
        ref = None
        for inline in (1, 0):
            s.set_option(_capi.OPT_INLINE_SEED, inline)
            for _ in range(5): out = s.search_device(q, 10)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): out = s.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            st = s.stats(); lab = out[1].clone()
            if ref is None: ref = lab
            print(json.dumps({"rows": rows, "B": B, "inline": inline, "launches": st["launches"], "splits": st["n_splits"], "path": st["path"],
                              "ms": round(e0.elapsed_time(e1) / 20, 4), "score_ms": round(st["score_ms"], 4), "same": bool(torch.equal(ref, lab))}), flush=True)
    s.close()
