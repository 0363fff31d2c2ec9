mkdir -p gpurun_out
python scripts/dbg_multipass.py > gpurun_out/r2f_dbg.log 2>&1; grep -c "bad=0" gpurun_out/r2f_dbg.log; grep -v "bad=0" gpurun_out/r2f_dbg.log | cut -c1-200
python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r2f_pytest.log | cut -c1-220
python /dev/stdin 10000000 > gpurun_out/r2f_b1.log 2>&1 <<'P'
import sys; sys.path.insert(0, ".")
import torch, image_recommender_b200 as irb
from image_recommender_b200 import _capi
R = int(sys.argv[1])
s = irb.FlatShard([48,128,1792], R, device=0); s.fill_synthetic(R, total_rows=R)
for fused in (1, 0):
    s.set_option(_capi.OPT_FUSED_TAIL, fused)
    for nq in (1, 8, 32, 128):
        q = s.synth_queries_device(nq, total_rows=R)
        for _ in range(5): s.search_device(q, 10)
        torch.cuda.synchronize(); s.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): s.search_device(q, 10)
        e1.record(); torch.cuda.synchronize(); st = s.stats()
        print("rows", R, "fused", fused, "nq", nq, "ms", round(e0.elapsed_time(e1)/20, 4), "score", round(st["score_ms"], 4), "tail", round(st["tail_ms"], 4), "launches", st["launches"], "cands", st["n_candidates"], flush=True)
P
cat gpurun_out/r2f_b1.log
timeout 900 python scripts/exp_tn.py > gpurun_out/r2f_tn.log 2>&1; echo "tn rc $?"; cut -c1-260 gpurun_out/r2f_tn.log
