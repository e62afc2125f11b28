#!/usr/bin/env python
"""
DYNAMIC opcode histogram of one kernel of an ncu report (--set full): executed thread instructions per ray, by SASS
opcode, from `ncu -i REP --page source --csv --print-source sass`.

    python tools/ncu_opcodes.py gpurun_out/r02e_prof_cfg5_grid.ncu-rep 1e8 [title] > profiles/r02_dynamic_opcodes_cfg5_grid.md
"""
import collections
import csv
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def main():
    rep, rays = sys.argv[1], float(sys.argv[2])
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    column, ops = None, collections.Counter()
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "Address":
            column = row.index("Instructions Executed")
        elif column is not None and len(row) > column and row[0].startswith("0x"):
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", row[1])
            if m:
                ops[m.group(1)] += int(row[column])
    scale = 32.0 / rays
    total = sum(ops.values()) * scale
    fp64 = sum(ops[o] for o in FP64) * scale
    print(f"# executed instructions per ray by opcode: {title}\n")
    print(f"source report: `{rep}`; warp-level `Instructions Executed` x 32 / {rays:g} rays\n")
    print(f"total **{total:.1f}**, of which FP64 pipe ({', '.join(FP64)}) **{fp64:.1f}** ({100 * fp64 / total:.0f} %)\n")
    print("| opcode | per ray |\n|---|---|")
    for op, n in ops.most_common():
        if n * scale >= 0.05:
            print(f"| `{op}` | {n * scale:.1f} |")


if __name__ == "__main__":
    main()
