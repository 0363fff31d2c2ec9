"""GPU debug: two indexes searched concurrently on two streams — per-iteration wall time, per index size and
with / without the in-kernel seeding (whose grid barrier assumes co-resident CTAs)."""
import sys
import time
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
DIMS = [48, 128, 1792]
for n, nq in ((300_000, 32), (2_000_000, 32), (2_000_000, 128), (2_000_000, 1)):
    for inline in (1, 0):
        shards, qs = [], []
        for s in range(2):
            ix = irb.FlatShard(DIMS, n, device=0)
            ix.fill_synthetic(n, total_rows=n, seed=0xC0FFEE + s)
            ix.set_option(_capi.OPT_INLINE_SEED, inline)
            shards.append(ix)
            qs.append(ix.synth_queries_device(nq, total_rows=n, seed=0xC0FFEE + s))
            ix.search_device(qs[-1], 10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            for ix, q in zip(shards, qs):
                ix.search_device(q, 10)
        torch.cuda.synchronize()
        serial = (time.perf_counter() - t0) / 5
        st = shards[0].stats()
        streams = [torch.cuda.Stream(device=0) for _ in range(2)]
        times = []
        for _ in range(12):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for ix, q, s_ in zip(shards, qs, streams):
                with torch.cuda.stream(s_):
                    ix.search_device(q, 10)
            torch.cuda.synchronize()
            times.append(round((time.perf_counter() - t0) * 1e3, 3))
        st2 = [ix.stats() for ix in shards]
        print(f"n={n} nq={nq} inline={inline} path={st['path']} splits={st['n_splits']} launches={st['launches']} "
              f"serial_ms={serial * 1e3:.3f} concurrent_ms={times} score_ms={[round(s['score_ms'], 3) for s in st2]} "
              f"tail_ms={[round(s['tail_ms'], 3) for s in st2]}", flush=True)
        for ix in shards:
            ix.close()
