"""GPU experiment: K-pack under a power-capped SM clock — bulk-copy kernel vs the two-pass kernel
(B2K_PACK_NO_BULK=1 in the environment selects the latter).  A sustained GEMM phase drives the clocks down,
then 56 pack launches are timed one by one."""
import json, os, statistics, sys, time
sys.path.insert(0, ".")
import torch
import image_recommender_b200 as irb
DIMS = [48, 128, 1792]; D = sum(DIMS); Dp = 1984
dev = torch.device("cuda", 0)
n_pack, per_round = 131072, 8
scratch = irb.FlatShard(DIMS, n_pack * per_round, device=0)
tabs = [torch.randn((n_pack, d), device=dev, dtype=torch.float32) for d in DIMS]
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
import pynvml; pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for phase in ("cool", "after 3 s of GEMM", "interleaved with GEMM"):
    times, clk = [], []
    for r in range(7):
        if phase != "cool":
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < (3.0 if (r == 0 or phase.startswith("inter")) and phase != "cool" else 0.0):
                for _ in range(20): torch.matmul(a, b)
                torch.cuda.synchronize()
        scratch.reset()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(per_round)]
        for e0, e1 in evs:
            e0.record(); scratch.add_tables_device(tabs); e1.record()
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        torch.cuda.synchronize()
        times += [e0.elapsed_time(e1) for e0, e1 in evs]
    # yardstick in the same phase: a device-to-device copy of the bytes the pack moves (read 7.9 KB, write 11.8 KB per row
    # is not a 50/50 copy, but it shows what the memory system delivers at this clock)
    src = torch.empty(n_pack * 2560, device=dev, dtype=torch.float32); dst = torch.empty_like(src)
    ce = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    if phase != "cool":
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 3.0:
            for _ in range(20): torch.matmul(a, b)
            torch.cuda.synchronize()
    for e0, e1 in ce:
        e0.record(); dst.copy_(src); e1.record()
    torch.cuda.synchronize()
    copy_gbps = 2 * src.numel() * 4 / statistics.median([e0.elapsed_time(e1) for e0, e1 in ce]) / 1e6
    del src, dst
    ms = statistics.median(times)
    by = n_pack * (4.0 * D + 4.0 * D + 2.0 * Dp + 4.0)
    print(json.dumps({"kernel": "two-pass (pack_rows_kernel)" if os.environ.get("B2K_PACK_NO_BULK") else "bulk (pack_rows_bulk_kernel)", "phase": phase, "median_ms": round(ms, 4),
                      "min_ms": round(min(times), 4), "max_ms": round(max(times), 4), "GBps": round(by / ms / 1e6, 1), "copy_GBps_same_phase": round(copy_gbps, 1), "sm_mhz": clk}), flush=True)
