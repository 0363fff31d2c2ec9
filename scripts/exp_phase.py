"""Debug build only (B2K_NVCC_EXTRA=-DB2K_PHASE_TIMERS): phase times of the fused tail kernel (CTA 0 = rank 0 of
query 0's cluster) at small batches, beside the tail's device time from the stats."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
lib = _capi.load_library()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
s = irb.FlatShard([48, 128, 1792], rows, device=0)
s.fill_synthetic(rows, total_rows=rows)
for B in (1, 128, 4096):
    q = s.synth_queries_device(B, total_rows=rows)
    for _ in range(4):
        s.search_device(q, 10)
        torch.cuda.synchronize()
        st = s.stats()
        t = (C.c_uint64 * 16)()
        lib.b2k_debug_phase_times(t)
        print(B, [int(t[i + 1] - t[i]) for i in range(5)], "ns: select, cluster barrier, re-rank, cluster barrier, finalize;",
              [int(t[6] - t[0]), int(t[7] - t[6]), int(t[8] - t[7]), int(t[1] - t[8])], "ns inside select: list load, top-k, tighten, emit;",
              "tail_ms", round(st["tail_ms"], 4), "cands", st["n_candidates"])
