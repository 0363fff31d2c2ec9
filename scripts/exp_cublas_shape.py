"""GPU experiment: what cuBLAS reaches on THIS contraction's shape (scores only, no top-k, output written to
HBM), next to the 8192^3 GEMM the bf16 peak in MEASURED_PEAKS.json comes from — on the same box, back to back
with this repo's fused kernel.  Library GEMM used as a yardstick only (never on the product path)."""
import json, sys, time
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb

dev = torch.device("cuda", 0)
def timed(fn, reps, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {}
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
ms = timed(lambda: torch.matmul(a, b), 20)
res["cublas_8192^3_tflops_burst"] = round(2 * 8192 ** 3 / ms / 1e9, 1)
t0 = time.time(); n = 0
while time.time() - t0 < 4.0:
    ms = timed(lambda: torch.matmul(a, b), 20, warm=0); n += 1
res["cublas_8192^3_tflops_sustained"] = round(2 * 8192 ** 3 / ms / 1e9, 1)
del a, b
B, Dp, D = 4096, 1984, 1968
q = torch.randn(B, Dp, device=dev, dtype=torch.bfloat16)
for n_rows in (262144, 1048576):
    db = torch.randn(n_rows, Dp, device=dev, dtype=torch.bfloat16)
    out = torch.empty(B, n_rows, device=dev, dtype=torch.bfloat16)
    ms = timed(lambda: torch.matmul(q, db.t(), out=out), 10)
    res[f"cublas_scores_only_{B}x{n_rows}x{Dp}_tflops"] = round(2.0 * B * n_rows * D / ms / 1e9, 1)
    t0 = time.time()
    while time.time() - t0 < 3.0:
        ms = timed(lambda: torch.matmul(q, db.t(), out=out), 10, warm=0)
    res[f"cublas_scores_only_{B}x{n_rows}x{Dp}_tflops_sustained"] = round(2.0 * B * n_rows * D / ms / 1e9, 1)
    del db, out
# this repo's fused kernel on the same box, 2 M rows (same per-CTA work profile as the bench's 10 M: >= 128 tiles per split)
rows = 5_000_000
s = irb.FlatShard([48, 128, 1792], rows, device=0)
s.fill_synthetic(rows, total_rows=rows)
qq = s.synth_queries_device(B, total_rows=rows)
for _ in range(3): s.search_device(qq, 10)
sc = []
for _ in range(8):
    s.search_device(qq, 10); sc.append(s.stats()["score_ms"])
res["b2k_fused_score_topk_tflops_5M_rows"] = round(2.0 * B * rows * D / (sum(sc) / len(sc)) / 1e9, 1)
print(json.dumps(res))
