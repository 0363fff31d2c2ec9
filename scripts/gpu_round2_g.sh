mkdir -p gpurun_out
python - > gpurun_out/r2g_occ.log 2>&1 <<'P'
import sys; sys.path.insert(0, ".")
import ctypes as C, torch
from image_recommender_b200 import _capi
lib = _capi.load_library()
torch.cuda.init(); torch.zeros(1, device="cuda")
for name in ("_ZN3b2k23score_tc_max_coresidentEi", "_ZN3b2k24score_tc2_max_coresidentEi"):
    f = getattr(lib, name); f.restype = C.c_int; f.argtypes = [C.c_int]
    print(name, f(148))
P
cat gpurun_out/r2g_occ.log
python scripts/exp_pack.py > gpurun_out/r2g_pack.log 2>&1; B2K_PACK_NO_BULK=1 python scripts/exp_pack.py >> gpurun_out/r2g_pack.log 2>&1; cut -c1-200 gpurun_out/r2g_pack.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pack or synthetic or save or load or ingest or end_to_end" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2g_pytest.log | cut -c1-200
cat > /tmp/tnprof.py <<'P'
import sys; sys.path.insert(0, ".")
import torch, image_recommender_b200 as irb
from image_recommender_b200 import _capi
R = 10000000
s = irb.FlatShard([48,128,1792], R, device=0); s.fill_synthetic(R, total_rows=R)
s.set_option(_capi.OPT_TN, 1)
q = s.synth_queries_device(int(sys.argv[1]), total_rows=R)
for _ in range(3): s.search_device(q, 10)
torch.cuda.synchronize()
P
for b in 128 160; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tn_kernel -s 1 -c 1 -o gpurun_out/prof_tn_$b python /tmp/tnprof.py $b > gpurun_out/ncu_tn_$b.log 2>&1; echo "ncu $b rc $?"
done
