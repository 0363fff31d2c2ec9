"""Debug build only (B2K_NVCC_EXTRA=-DB2K_PHASE_TIMERS): phase times of the in-kernel seeding, CTA 0 and the last CTA."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
lib = _capi.load_library()
rows = 1_250_000
s = irb.FlatShard([48, 128, 1792], rows, device=0)
s.fill_synthetic(rows, total_rows=rows)
names = ["start->tile0 drained", "list store", "barrier 1", "handler", "barrier 2", "rest of the tiles", "final list store"]
for B in (4, 8, 128):
    q = s.synth_queries_device(B, total_rows=rows)
    for _ in range(4):
        s.search_device(q, 10)
        torch.cuda.synchronize()
    t = (C.c_uint64 * 16)()
    lib.b2k_debug_tc_phase_times(t)
    st = s.stats()
    for cta in (0, 1):
        d = [int(t[cta * 8 + i + 1]) - int(t[cta * 8 + i]) for i in range(7)]
        print(B, "cta", "first" if cta == 0 else "last", dict(zip(names, d)), "ns; score_ms", round(st["score_ms"], 4), flush=True)
