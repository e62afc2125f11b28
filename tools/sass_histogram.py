#!/usr/bin/env python
"""
SASS opcode histograms of the hot kernels, made WITHOUT a GPU (nvcc + cuobjdump):

* the run-time specialised kernels (``optk_jit_kernel``) for the BASELINE systems, compiled offline by
  ``tools/jit_offline.py`` exactly as jit.cu hands them to NVRTC;
* the bulk-copy pipeline ``trace_kernel_tma`` from the built ``liboptk.so`` (``UBLKCP`` = cp.async.bulk,
  ``SYNCS`` = mbarrier operations).

    python tools/sass_histogram.py > profiles/r02_sass_histograms.md
"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def histogram(sass: str) -> collections.Counter:
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    return ops


def table(title: str, ops: collections.Counter, note: str = "") -> str:
    total = sum(ops.values())
    rows = [f"### {title}", "", f"{total} SASS instructions (static). {note}", "", "| opcode | count |", "|---|---|"]
    rows += [f"| `{op}` | {n} |" for op, n in ops.most_common()]
    return "\n".join(rows) + "\n"


def jit(cfg: str, mode: str) -> tuple:
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "jit_offline.py"), cfg, mode, "--sass"],
                         capture_output=True, text=True).stdout
    sass = out.split("=== SASS ===", 1)[1] if "=== SASS ===" in out else ""
    regs = re.findall(r"Used (\d+) registers", out)
    return histogram(sass), (regs[-1] if regs else "?")


def main():
    print("# SASS opcode histograms (round 2; static counts from `cuobjdump -sass`, no GPU needed)\n")
    print("Made by `tools/sass_histogram.py`.  `DFMA/DMUL/DADD/DSETP` run on the FP64 pipe, `MUFU` on the XU pipe,")
    print("`SHFL/REDUX/VOTE` are the warp aggregation of the detector binning, `RED`/`ATOMG` the global reductions,")
    print("`UBLKCP` is `cp.async.bulk` and `SYNCS` the mbarrier traffic of the bulk-copy pipeline.  There is no")
    print("`UTCMMA` / `HMMA`: nothing on this path is a contraction.\n")
    for cfg, mode, what in (
        ("cfg2", "dense", "cfg 2, dense SoA in -> dense SoA out (the bench `value` kernel)"),
        ("cfg2", "image", "cfg 2, fused trace + detector binning from broadcast grids (the bench `e2e` kernel)"),
        ("cfg2", "grid", "cfg 2, generate (Philox) + trace + bin (`optk_trace_grid`)"),
        ("cfg3", "dense", "cfg 3 (toroid Newton, polynomial rulings, octagon), dense"),
        ("cfg3", "grid", "cfg 3, generate + trace + bin (the cfg3_strong kernel)"),
        ("cfg1", "grid", "cfg 1 / cfg 5 walk, generate + trace + bin (the cfg5_strong kernel)"),
    ):
        ops, regs = jit(cfg, mode)
        print(table(f"`optk_jit_kernel`: {what}", ops, f"{regs} registers per thread, two rays per thread."))
    lib = ROOT / "optika_b200" / "liboptk.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    for block in blocks:
        name = block.split("\n", 1)[0]
        if "trace_kernel_tma" in name:
            print(table(f"`{name.strip()[:90]}` (bulk-copy pipeline, from liboptk.so)", histogram("\n" + block)))
            break


if __name__ == "__main__":
    main()
