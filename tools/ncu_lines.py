#!/usr/bin/env python
"""
Executed warp instructions per SOURCE LINE of one kernel of an ncu report (--set full, -lineinfo):

    python tools/ncu_lines.py gpurun_out/r02b_prof_cfg3_grid.ncu-rep [rays] [top]

Reads `ncu -i REP --page source --csv --print-source cuda,sass`, sums "Instructions Executed" per (file, line),
prints the heaviest lines with the text of the line from optika_b200/csrc, and the total per file.  With `rays`
the counts are per ray (warp instructions x 32 / rays -- the unit of profiles/r02_instruction_counts.json is
warp-level instructions per ray x 32 ... i.e. thread instructions per ray at full occupancy of the warp).
"""
import collections
import csv
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def main():
    rep = sys.argv[1]
    rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(
        ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True
    ).stdout
    per_line = collections.Counter()
    fp64 = collections.Counter()
    path, column, line = None, None, None
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == "File Path":
            path = pathlib.Path(row[1]).name
            continue
        if row[0] == "Line No":
            column = row.index("Instructions Executed")
            continue
        if column is None or len(row) <= column:
            continue
        if row[0].strip().isdigit():
            line = int(row[0])
            per_line[(path, line)] += int(row[column])
        elif line is not None and row[3].strip()[:2] in ("DF", "DM", "DA", "DS"):
            pass
    scale = 32.0 / rays if rays else 1.0
    total = sum(per_line.values())
    print(f"total {total * scale:.1f}")
    files = collections.Counter()
    for (p, _), v in per_line.items():
        files[p] += v
    for p, v in files.most_common():
        print(f"  {p}: {v * scale:.1f}")
    for (p, ln), v in per_line.most_common(top):
        text = ""
        f = ROOT / "optika_b200" / "csrc" / p
        if f.exists():
            lines = f.read_text().splitlines()
            if ln - 1 < len(lines):
                text = lines[ln - 1].strip()
        print(f"{v * scale:9.2f}  {p}:{ln}  {text[:110]}")


if __name__ == "__main__":
    main()
