"""torchrun check of the row-sharded host API (image_recommender_b200.sharded.ShardedIndex):
every rank loads its row range of ONE index file; results must equal a single-GPU load bit for bit,
through the NVLink peer exchange and through the NCCL all-gather path.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded.py"""
import json, os, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist
import image_recommender_b200 as irb
from image_recommender_b200.sharded import ShardedIndex

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DIMS, N = [48, 128, 1792], 150_001
path = Path(tempfile.gettempdir()) / "b2k_check_sharded.faiss"
whole = None
if rank == 0:
    whole = irb.FlatShard(DIMS, N, device=local)
    whole.fill_synthetic(N, total_rows=N, n_clusters=64)
    whole.save(path, np.arange(N, dtype=np.int64) + 1)
dist.barrier()
probe = irb.FlatShard(DIMS, 8, device=local)
res = {"world": world, "rows": N}
for peer in (True, False):
    ix = ShardedIndex.load(path, device=local, peer=peer)
    assert ix.ntotal == N
    ok = True
    for nq in (1, 7, 300, 5000):
        q = probe.synth_queries_device(nq, total_rows=N, n_clusters=64, qseed=nq).cpu().numpy()
        d, l = ix.search(q, 10)
        if rank == 0:
            wd, wl = whole.search(q, 10)
            ok = ok and np.array_equal(l, wl) and np.array_equal(d.view(np.uint32), wd.view(np.uint32))
        # every rank holds the same answer
        t = torch.from_numpy(l).cuda()
        ref = t.clone(); dist.broadcast(ref, 0)
        ok = ok and bool(torch.equal(t, ref))
    res["peer" if peer else "nccl"] = ok
    ix.close()
    dist.barrier()
flags = torch.tensor([int(res["peer"]), int(res["nccl"])], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    res["peer"], res["nccl"] = bool(flags[0].item()), bool(flags[1].item())
    print(json.dumps(res), flush=True)
    path.unlink(missing_ok=True)
dist.destroy_process_group()
sys.exit(0 if flags.min().item() == 1 else 1)
