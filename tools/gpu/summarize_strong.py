#!/usr/bin/env python
"""Print the strong-scaling objects of bench lines: python tools/gpu/summarize_strong.py file.json ..."""
import json
import sys

KEYS = ("transport", "ms_total", "ms_trace", "ms_reduce", "ms_d2h", "intercepts_per_s", "efficiency_vs_n1", "counts_equal_n1")
for path in sys.argv[1:]:
    try:
        for line in open(path):
            if line.strip().startswith("{"):
                config = json.loads(line)["config"]
                for name in ("cfg3_strong", "cfg5_strong"):
                    if name in config:
                        v = config[name]
                        row = {k: (round(v[k], 3) if isinstance(v.get(k), float) else v.get(k)) for k in KEYS}
                        n1 = v.get("n1_same_run")
                        print(path, name, row, "n1 ms", round(n1["ms_total"], 2) if n1 else None)
                        for key, label in (("two_planes", "two planes"), ("one_plane", "one plane")):
                            two = v.get(key)
                            if two:
                                print(f"    {label}:", {k: (round(two[k], 3) if isinstance(two.get(k), float) else two.get(k))
                                                        for k in ("ms_total", "ms_trace", "ms_reduce", "ms_d2h", "efficiency_vs_n1")})
    except Exception as e:  # noqa: BLE001
        print(path, e)
