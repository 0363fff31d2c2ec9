"""GPU experiment: sampling pass of the threshold seeding on one wave of long strided CTAs (B2K_OPT_SAMPLE_WAVE)
vs on the main pass's grid (every split its first tiles).  The two variants alternate search by search (the
chip's power / thermal state drifts by several per cent within seconds); medians over the rounds."""
import sys, json, statistics
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
OPT = int(sys.argv[1]) if len(sys.argv) > 1 else _capi.OPT_SAMPLE_WAVE
VALS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
cases = [([48, 128, 1792], 10_000_000, (160, 256, 384, 512, 1024, 4096)), ([48, 128, 1792], 1_250_000, (256, 384, 1024, 4096)),
         ([1792], 1_000_000, (1000,))]
if len(sys.argv) > 3 and sys.argv[3] == "short":
    cases = [([48, 128, 1792], 10_000_000, (512, 4096)), ([48, 128, 1792], 1_250_000, (4096,))]
if len(sys.argv) > 3 and sys.argv[3] == "tail":
    cases = [([48, 128, 1792], 1_250_000, (32, 128, 256, 512, 1024, 2048, 4096)), ([48, 128, 1792], 10_000_000, (128, 1024, 4096))]
if len(sys.argv) > 3 and sys.argv[3] == "big":
    cases = [([48, 128, 1792], 1_250_000, (512, 1024, 2048, 4096)), ([48, 128, 1792], 10_000_000, (1024, 4096))]
if len(sys.argv) > 3 and sys.argv[3] == "mid":
    cases = [([48, 128, 1792], 10_000_000, (224, 240, 256)), ([48, 128, 1792], 5_000_000, (224, 256))]
if len(sys.argv) > 3 and sys.argv[3] == "tnshort":
    cases = [([48, 128, 1792], 1_250_000, (130, 160, 192, 208)), ([48, 128, 1792], 2_500_000, (130, 160, 192, 208)),
             ([48, 128, 1792], 5_000_000, (130, 160, 192, 208))]
if len(sys.argv) > 3 and sys.argv[3] == "small":
    cases = [([48, 128, 1792], 1_250_000, (16, 32, 64, 128, 256, 512)), ([48, 128, 1792], 10_000_000, (32, 128, 256, 512))]
if len(sys.argv) > 3 and sys.argv[3] == "odd":
    cases = [([48, 128, 1792], 10_000_000, (300, 384, 600, 640, 860, 896)), ([48, 128, 1792], 1_250_000, (384, 640, 896))]
if len(sys.argv) > 3 and sys.argv[3] == "shards":
    cases = [([48, 128, 1792], 5_000_000, (512, 4096)), ([48, 128, 1792], 2_500_000, (512, 4096)), ([48, 128, 1792], 1_250_000, (512, 4096))]
for dims, rows, batches in cases:
    s = irb.FlatShard(dims, rows, device=0)
    s.fill_synthetic(rows, total_rows=rows)
    D = sum(dims)
    for B in batches:
        q = s.synth_queries_device(B, total_rows=rows)
        ms = {v: [] for v in VALS}
        sc = {v: [] for v in VALS}
        tl = {v: [] for v in VALS}
        labs = {}
        rounds = 15 if B <= 1024 else 7
        for r in range(rounds + 1):
            for v in VALS:
                s.set_option(OPT, v)
                s.search_device(q, 10)
                torch.cuda.synchronize(); s.stats()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2): out = s.search_device(q, 10)
                e1.record(); torch.cuda.synchronize()
                st = s.stats()
                if r:
                    ms[v].append(e0.elapsed_time(e1) / 2); sc[v].append(st["score_ms"]); tl[v].append(st["tail_ms"])
                labs[v] = out[1].clone()
        print(json.dumps({"dims": dims, "rows": rows, "B": B, "opt": OPT, "path": st["path"], "launches": st["launches"],
                          "median_ms": {v: round(statistics.median(ms[v]), 3) for v in VALS},
                          "min_ms": {v: round(min(ms[v]), 3) for v in VALS},
                          "median_score_ms": {v: round(statistics.median(sc[v]), 3) for v in VALS},
                          "median_tail_ms": {v: round(statistics.median(tl[v]), 3) for v in VALS},
                          "same": all(bool(torch.equal(labs[VALS[0]], labs[v])) for v in VALS)}), flush=True)
    s.close()
