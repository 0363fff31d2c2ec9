"""GPU debug (run under ncu --metrics gpu__time_duration.sum): one search per tail variant at batch 4096."""
import sys
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
s = irb.FlatShard([48, 128, 1792], rows, device=0)
s.fill_synthetic(rows, total_rows=rows)
q = s.synth_queries_device(4096, total_rows=rows)
for v in (0, 1, 2):
    s.set_option(_capi.OPT_FUSED_TAIL, v)
    for _ in range(2):
        s.search_device(q, 10)
    torch.cuda.synchronize()
