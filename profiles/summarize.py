#!/usr/bin/env python
"""
Turn an Nsight Compute report (`.ncu-rep`, brought back from the GPU box in
`gpurun_out/`) into the text summary that is committed under `profiles/`.

    python profiles/summarize.py gpurun_out/prof_trace_r1.ncu-rep --rays 1e8 > profiles/r01_trace_kernel.md

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv
--print-source cuda,sass` (needs the kernels to be built with -lineinfo).
"""

import argparse
import collections
import csv
import io
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs / SM (register limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active, % of peak"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe, % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe, %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "ADU pipe (indexed constant loads), %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (MUFU), %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "global RED requests"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 RED sectors"),
]


def ncu(report, *args):
    out = subprocess.run(["ncu", "-i", report, *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--rays", type=float, default=None, help="rays per launch (for per-ray figures)")
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()

    raw = ncu(args.report, "--page", "raw", "--csv")
    header, units, values = raw[0], raw[1], raw[2]
    metric = dict(zip(header, values))
    unit = dict(zip(header, units))
    print(f"# ncu summary: `{metric.get('Kernel Name', '?')}`\n")
    print(f"source report: `{args.report}` (`ncu --set full --clock-control none --import-source on`)\n")
    print("| metric | value |\n|---|---|")
    for key, label in KEYS:
        if key in metric:
            print(f"| {label} (`{key}`) | {metric[key]} {unit.get(key, '')} |")
    if args.rays and "smsp__inst_executed.sum" in metric:
        per_ray = float(metric["smsp__inst_executed.sum"]) * 32 / args.rays
        print(f"| instructions per ray | {per_ray:.0f} |")
    if "dram__bytes_read.sum" in metric and args.rays:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        total = sum(
            float(metric[k]) * scale.get(unit.get(k, "byte"), 1.0)
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")
        )
        print(f"| DRAM traffic per ray | {total / args.rays:.1f} B |")
    print("\n## warp stall reasons (stalled warps per issued instruction)\n")
    print("| reason | value |\n|---|---|")
    stalls = []
    for h, v in metric.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            name = h[len("smsp__average_warps_issue_stalled_") : -len("_per_issue_active.ratio")]
            try:
                stalls.append((float(v), name))
            except ValueError:
                pass
    for v, name in sorted(stalls, reverse=True):
        if v > 0.01:
            print(f"| {name} | {v:.2f} |")

    rows = ncu(args.report, "--page", "source", "--csv", "--print-source", "cuda,sass")
    per, samp, text = collections.Counter(), collections.Counter(), {}
    hdr, cur = None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 3 and r[0] == "Line No":
            hdr = r
            ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ia or r[0] == "":
            continue
        try:
            n, s = int(r[ia]), int(r[isamp])
        except ValueError:
            continue
        key = (cur, int(r[0]))
        per[key] += n
        samp[key] += s
        text[key] = r[1].strip()
    total_samples = sum(samp.values()) or 1
    print(f"\n## top {args.top} source lines by stall samples (of {total_samples})\n")
    print("| file:line | samples % | warp instr | source |\n|---|---|---|---|")
    for key, n in samp.most_common(args.top):
        src = text[key].replace("|", "\\|")[:100]
        print(f"| {key[0]}:{key[1]} | {100 * n / total_samples:.1f} | {per[key]} | `{src}` |")


if __name__ == "__main__":
    main()
