"""GPU experiment: K-scan vs K-score (tcgen05) across row widths at small batches."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi

def run(dims, rows, batches=(1, 4, 64), k=10):
    D = sum(dims)
    s = irb.FlatShard(dims, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    for B in batches:
        q = s.synth_queries_device(B, total_rows=rows)
        ref = None
        for path in (1, 2):
            if path == 1 and B > 4: continue
            s.set_option(_capi.OPT_PATH, path)
            try:
                for _ in range(3): out = s.search_device(q, k)
            except Exception as e:
                print(json.dumps({"dims": dims, "B": B, "path": path, "error": str(e)[:80]})); continue
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out = s.search_device(q, k)
            e1.record(); torch.cuda.synchronize()
            st = s.stats(); lab = out[1].clone()
            if ref is None: ref = lab
            print(json.dumps({"dims": dims, "rows": rows, "B": B, "path": st["path"], "ms": round(e0.elapsed_time(e1) / 10, 4),
                              "score_ms": round(st["score_ms"], 4), "tail_ms": round(st["tail_ms"], 4),
                              "hbm_frac": round(2.0 * rows * D / (st["score_ms"] * 1e-3) / 1e9 / 6558.1, 3),
                              "unc": st["n_uncertified"], "sat": st["n_saturated"], "same": bool(torch.equal(ref, lab))}), flush=True)
    s.close()

run([48], 10_000_000)
run([128], 5_000_000)
run([48, 128], 5_000_000)
run([512], 4_000_000)
run([1024], 4_000_000)
run([1792], 2_000_000)
