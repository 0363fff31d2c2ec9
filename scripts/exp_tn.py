"""GPU experiment: transposed K-score kernel (B2K_OPT_TN) vs the M = queries kernels."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
cases = [([48, 128, 1792], 10_000_000, (1, 8, 32, 64, 128, 130, 160, 192, 224, 240, 256)),
         ([48], 10_000_000, (1, 8, 32, 64)), ([128], 5_000_000, (1, 8, 32, 64))]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    cases = [([48, 128, 1792], 1_250_000, (1, 32, 130, 160, 192, 256))]
for dims, rows, batches in cases:
    s = irb.FlatShard(dims, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    D = sum(dims)
    for B in batches:
        q = s.synth_queries_device(B, total_rows=rows)
        ref = None
        for tn in (0, 1):
            s.set_option(_capi.OPT_TN, tn)
            for _ in range(3): out = s.search_device(q, 10)
            torch.cuda.synchronize(); s.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out = s.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            st = s.stats(); lab = out[1].clone()
            if ref is None: ref = lab
            ms = e0.elapsed_time(e1) / 10
            print(json.dumps({"dims": dims, "rows": rows, "B": B, "tn": tn, "path": st["path"], "ms": round(ms, 3), "score_ms": round(st["score_ms"], 3),
                              "tail_ms": round(st["tail_ms"], 3), "qps": round(B / ms * 1e3), "hbm_GBps": round(2.0 * rows * D / st["score_ms"] / 1e6),
                              "launches": st["launches"], "unc": st["n_uncertified"], "sat": st["n_saturated"], "same": bool(torch.equal(ref, lab))}), flush=True)
    s.close()
