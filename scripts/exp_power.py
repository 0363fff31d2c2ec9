"""GPU experiment: steady-state behaviour of the scoring kernels under the power cap — per-search kernel time,
SM clock and board power while a search is repeated back to back."""
import sys, json, time, threading
sys.path.insert(0, ".")
import torch, pynvml
import image_recommender_b200 as irb
from image_recommender_b200 import _capi
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
R = 10_000_000
s = irb.FlatShard([48, 128, 1792], R, device=0); s.fill_synthetic(R, total_rows=R)
def run(B, tn, iters=60):
    s.set_option(_capi.OPT_TN, tn)
    q = s.synth_queries_device(B, total_rows=R)
    torch.cuda.synchronize(); time.sleep(2.0)          # start from an idle, cool chip
    clk, pw, stop = [], [], False
    def sample():
        while not stop:
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0); time.sleep(0.01)
    th = threading.Thread(target=sample); th.start()
    per = []
    for i in range(iters):
        s.search_device(q, 10)
        if i % 10 == 9:
            st = s.stats(); per.append(round(st["score_ms"], 3))
    stop = True; th.join()
    st = s.stats()
    print(json.dumps({"B": B, "tn": tn, "path": st["path"], "score_ms_per_10": per, "sm_mhz_first_last": [clk[0], clk[len(clk)//2], clk[-1]], "power_w_max": round(max(pw)), "power_w_last": round(pw[-1])}), flush=True)
for B in (128, 160, 256):
    for tn in (0, 1):
        run(B, tn)
