#!/usr/bin/env python
"""Source-level stall summary of one ncu report: the SASS instructions that hold the most warp-stall
samples, with the warp role they belong to (the kernels are warp-specialised, so an address range = a role).

    python scripts/ncu_stalls.py gpurun_out/prof_mid_256.ncu-rep [top_n]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"]) for r in rows) or 1
print(f"# {rep}: {len(rows)} SASS instructions, {tot} samples")
reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
print("# whole kernel by reason:", ", ".join(f"{k[6:]} {100 * sum(int(r[k]) for r in rows) / tot:.1f}%" for k in
      sorted(reasons, key=lambda k: -sum(int(r[k]) for r in rows))[:8]))
idx = {id(r): i for i, r in enumerate(rows)}
for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:top]:
    i = idx[id(r)]
    why = sorted(((int(r[k]), k[6:]) for k in reasons), reverse=True)[:2]
    print(f"{100 * int(r['# Samples']) / tot:5.1f}%  #{i:5d}  {r['Source'].strip()[:70]:70s}  {why[0][1]} {why[0][0]}, {why[1][1]} {why[1][0]}  exec={r['Instructions Executed']}")
