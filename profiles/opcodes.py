#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode from an Nsight Compute report:
    python profiles/opcodes.py gpurun_out/prof.ncu-rep --rays 1e8
"""
import argparse, collections, csv, io, re, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--rays", type=float, default=1e8)
ap.add_argument("--top", type=int, default=30)
a = ap.parse_args()
out = subprocess.run(
    ["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True
).stdout
hdr, ops, tot = None, collections.Counter(), 0
for r in csv.reader(io.StringIO(out)):
    if len(r) > 3 and "Source" in r and "Instructions Executed" in r:
        hdr, isrc, ia = r, r.index("Source"), r.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= ia:
        continue
    try:
        n = int(r[ia])
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc].strip())
    if m:
        ops[m.group(2).split(".")[0]] += n
        tot += n
print(f"thread instructions per ray: {tot * 32 / a.rays:.0f}")
for op, n in ops.most_common(a.top):
    print(f"{op:12s} {n * 32 / a.rays:8.1f}")
