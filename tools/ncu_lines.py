#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep
(needs -lineinfo at compile time and --import-source on at capture).
Usage: python tools/ncu_lines.py file.ncu-rep <kernel-regex> [which-instance] [top-n]"""
import csv
import io
import os
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
k, fname, agg, fn_seen = 0, "", {}, {}
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        if fn_seen.get("last") != r[1]:
            fn_seen["last"] = r[1]
            k = fn_seen.setdefault(r[1], len(fn_seen))
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if k != which or not r[0].isdigit():
        continue
    try:
        inst, smp, thr = int(r[7]), int(r[6]), int(r[8])
    except (ValueError, IndexError):
        continue
    if inst == 0 and smp == 0:
        continue
    a = agg.setdefault((fname, int(r[0]), r[1].strip()[:100]), [0, 0, 0])
    a[0] += inst
    a[1] += smp
    a[2] += thr
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot}, stall samples {tots}")
for (f, line, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot * 100:5.1f}% inst {a[1] / tots * 100:5.1f}% smp lanes {a[2] / max(1, a[0]):5.1f} | {f}:{line} {src}")
