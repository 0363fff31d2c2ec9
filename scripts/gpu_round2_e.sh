mkdir -p gpurun_out
python scripts/dbg_multipass.py > gpurun_out/r2e_dbg.log 2>&1; cat gpurun_out/r2e_dbg.log | cut -c1-300
cat > /tmp/b1.py <<'P'
import sys; sys.path.insert(0, ".")
import torch, image_recommender_b200 as irb
from image_recommender_b200 import _capi
R = int(sys.argv[1])
s = irb.FlatShard([48,128,1792], R, device=0); s.fill_synthetic(R, total_rows=R)
for fused in (1, 0):
    s.set_option(_capi.OPT_FUSED_TAIL, fused)
    for nq in (1, 8, 32):
        q = s.synth_queries_device(nq, total_rows=R)
        for _ in range(5): s.search_device(q, 10)
        torch.cuda.synchronize(); s.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): s.search_device(q, 10)
        e1.record(); torch.cuda.synchronize(); st = s.stats()
        print("rows", R, "fused", fused, "nq", nq, "ms", round(e0.elapsed_time(e1)/20, 4), "score", round(st["score_ms"], 4), "tail", round(st["tail_ms"], 4), "launches", st["launches"], "cands", st["n_candidates"], flush=True)
P
python /tmp/b1.py 10000000 > gpurun_out/r2e_b1.log 2>&1; cat gpurun_out/r2e_b1.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tail_kernel|select_kernel|rerank_kernel|finalize_kernel|collect|exact' -c 300 --csv --log-file gpurun_out/r2e_launches.csv python /tmp/b1.py 10000000 > /dev/null 2>&1; echo "ncu rc $?"
for v in two mb6 mb8; do B2K_PACK_VARIANT=$v python scripts/exp_pack.py 2>&1 | sed "s/^/$v /" >> gpurun_out/r2e_pack.log; done; cut -c1-170 gpurun_out/r2e_pack.log
python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc $?"; tail -12 gpurun_out/r2e_pytest.log | cut -c1-220
