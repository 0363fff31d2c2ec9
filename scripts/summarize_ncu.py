#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep and launches.csv into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py r01          # prefix for the output files
"""
import collections
import csv
import io
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_uniform"]


def launches(prefix):
    src = OUT / "launches.csv"
    if not src.exists():
        return
    lines = [l for l in src.read_text().splitlines() if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = ["kernel,launches,total_ms,avg_ms,share_pct"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k},{v[0]},{v[1]:.4f},{v[1] / v[0]:.5f},{100 * v[1] / tot:.2f}")
    (PROF / f"{prefix}_launches_by_kernel.csv").write_text("\n".join(out) + "\n")
    (PROF / f"{prefix}_launches_raw.csv").write_text("\n".join(lines) + "\n")
    print("\n".join(out))


def full(prefix, rep):
    p = OUT / f"{rep}.ncu-rep"
    if not p.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(p), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"# {rep}: {name}")
        for h, u, v in zip(hdr, units, vals):
            if any(h == k or h.startswith(k) for k in KEYS):
                out.append(f"{h},{v},{u}")
    det = subprocess.run(["ncu", "-i", str(p), "--page", "details"], capture_output=True, text=True).stdout
    keep = [l for l in det.splitlines() if l.strip() and not l.strip().startswith(("OPT", "INF", "---"))]
    (PROF / f"{prefix}_{rep}_metrics.csv").write_text("\n".join(out) + "\n")
    (PROF / f"{prefix}_{rep}_details.txt").write_text(det)
    print("\n".join(out[:40]))


if __name__ == "__main__":
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"
    PROF.mkdir(exist_ok=True)
    launches(prefix)
    for rep in sorted(p.stem for p in OUT.glob("*.ncu-rep")):
        full(prefix, rep)
