#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers
the roofline discussion needs.  Usage: python tools/ncu_summary.py file.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sectors.sum", "l2_sectors"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_%"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        out.append(f"### {name}")
        line = []
        for k, short in KEYS:
            if k in idx:
                line.append(f"{short}={d[idx[k]]} {units[idx[k]]}".strip())
        out.append(", ".join(line))
        st = []
        for h, i in idx.items():
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        out.append("stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))
        out.append("")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(text + "\n")


if __name__ == "__main__":
    main()
