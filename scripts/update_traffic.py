#!/usr/bin/env python
"""profiles/r01_traffic.json from the summarised ncu captures (dram bytes per launch of the scoring kernels on
the default bench workload; bench.py reports them as roofline.traffic when the workload matches)."""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"


def metrics(rep):
    out = {}
    p = ROOT / "profiles" / f"{prefix}_{rep}_metrics.csv"
    if not p.exists():
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "s": 1e3, "%": 1.0}
    for line in p.read_text().splitlines():
        m = re.match(r"([^,#]+),([-0-9.e+]+),(\S*)$", line)
        if m:
            out[m.group(1)] = float(m.group(2)) * scale.get(m.group(3), 1.0)
    return out


doc = {"_doc": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures of "
               "`python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (scripts/gpu_profile.sh); valid for that workload only"}
for kernel, rep, batch in (("score_tc2_kernel", "prof_score_tc2", 4096), ("score_tc_kernel", "prof_score_tc_b1", 1),
                           ("scan_bf16_kernel", "prof_scan", 1)):
    m = metrics(rep)
    if not m:
        continue
    doc[kernel] = {
        "rows_local": 10_000_000, "batch": batch,
        "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
        "duration_ms_under_ncu": m["gpu__time_duration.sum"],
        "tensor_active_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_pct_of_peak": m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "source": f"profiles/{prefix}_{rep}_metrics.csv",
    }
(ROOT / "profiles" / f"{prefix}_traffic.json").write_text(json.dumps(doc, indent=1) + "\n")
print(json.dumps(doc, indent=1))
