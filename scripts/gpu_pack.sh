mkdir -p gpurun_out
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
import image_recommender_b200 as irb
DIMS=[48,128,1792]; D=sum(DIMS)
dev=torch.device('cuda',0)
for n_pack in (16384, 131072, 524288):
    reps=6
    s=irb.FlatShard(DIMS, n_pack*(reps+2), device=0)
    tabs=[torch.randn((n_pack,d),device=dev) for d in DIMS]
    for _ in range(2): s.add_tables_device(tabs)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): s.add_tables_device(tabs)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    b=n_pack*(8.0*D+2*1984+4)
    print(n_pack, "ms",round(ms,4),"GB/s",round(b/ms/1e6,1), flush=True)
    s.close(); del tabs
PY
