"""A few batch-1 / batch-16 searches on an 8-GPU-sized shard (1.25 M combo rows): run under
`ncu --metrics gpu__time_duration.sum` to list the per-kernel device times of the pipeline."""
import sys
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
s = irb.FlatShard([48, 128, 1792], rows, device=0)
s.fill_synthetic(rows, total_rows=rows)
for B in (1, 16, 4096):
    q = s.synth_queries_device(B, total_rows=rows)
    for _ in range(4):
        out = s.search_device(q, 10)
    torch.cuda.synchronize()
    print(B, s.stats())
s.close()
