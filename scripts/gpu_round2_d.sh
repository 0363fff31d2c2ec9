mkdir -p gpurun_out
python scripts/dbg_tail.py > gpurun_out/r2d_dbg.log 2>&1; cat gpurun_out/r2d_dbg.log | cut -c1-300
cat > /tmp/b1.py <<'P'
import sys; sys.path.insert(0, ".")
import torch, image_recommender_b200 as irb
from image_recommender_b200 import _capi
s = irb.FlatShard([48,128,1792], 1250000, device=0); s.fill_synthetic(1250000, total_rows=1250000)
for fused in (1, 0):
    s.set_option(_capi.OPT_FUSED_TAIL, fused)
    for nq in (1, 32, 4096):
        q = s.synth_queries_device(nq, total_rows=1250000)
        for _ in range(5): s.search_device(q, 10)
        torch.cuda.synchronize(); s.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): s.search_device(q, 10)
        e1.record(); torch.cuda.synchronize(); st = s.stats()
        print("fused", fused, "nq", nq, "ms", round(e0.elapsed_time(e1)/20, 4), "score", round(st["score_ms"], 4), "tail", round(st["tail_ms"], 4), "launches", st["launches"], flush=True)
P
python /tmp/b1.py > gpurun_out/r2d_b1.log 2>&1; cat gpurun_out/r2d_b1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2d_launches.csv python /tmp/b1.py > /dev/null 2>&1; echo "ncu rc $?"
