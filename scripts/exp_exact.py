"""GPU experiment: cost of the exhaustive fallback (K-exact) and of K-collect on the combo workload."""
import sys, json
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
DIMS = [48, 128, 1792]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = irb.FlatShard(DIMS, rows, device=0)
s.fill_synthetic(rows, total_rows=rows)
for B, k in ((4, 10), (16, 10), (4, 100)):
    q = s.synth_queries_device(B, total_rows=rows)
    ref = s.search_device(q, k)[1].clone()
    s.set_option(_capi.OPT_FORCE_EXACT, 1)
    for _ in range(2): out = s.search_device(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = s.search_device(q, k)
    e1.record(); torch.cuda.synchronize()
    st = s.stats()
    s.set_option(_capi.OPT_FORCE_EXACT, 0)
    print(json.dumps({"rows": rows, "B": B, "k": k, "forced_exact_ms": round(e0.elapsed_time(e1) / 3, 3), "tail_ms": round(st["tail_ms"], 3),
                      "uncertified": st["n_uncertified"], "same": bool(torch.equal(ref, out[1])),
                      "fp32_gbs_per_pass": round(rows * 1968 * 4 / 1e9 / (st["tail_ms"] * 1e-3) * ((B + 3) // 4) * ((k + 31) // 32), 1)}), flush=True)
s.close()
