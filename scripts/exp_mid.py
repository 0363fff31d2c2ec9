"""GPU experiment: batches 129..512 — CTA-pair kernel vs single-CTA kernel with one wave of (qtile, split) CTAs."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
DIMS = [48, 128, 1792]; D = sum(DIMS)
for rows in (1_250_000, 10_000_000):
    s = irb.FlatShard(DIMS, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    for B in (160, 256, 320, 384, 512, 640):
        q = s.synth_queries_device(B, total_rows=rows)
        nqt = (B + 127) // 128
        ref = None
        for pair, sp in ((1, 0), (0, 148), (0, 148 // nqt), (0, 2 * (148 // nqt))):
            s.set_option(_capi.OPT_TC_PAIR, pair); s.set_option(_capi.OPT_SPLITS, sp)
            for _ in range(3): out = s.search_device(q, 10)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(6): out = s.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            st = s.stats(); lab = out[1].clone()
            if ref is None: ref = lab
            print(json.dumps({"rows": rows, "B": B, "pair": pair, "splits": st["n_splits"], "ms": round(e0.elapsed_time(e1) / 6, 3),
                              "score_ms": round(st["score_ms"], 3), "qps": round(B / (e0.elapsed_time(e1) / 6) * 1e3),
                              "unc": st["n_uncertified"], "same": bool(torch.equal(ref, lab))}), flush=True)
    s.close()
