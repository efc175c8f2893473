#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output into a small per-launch table (markdown).

usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py raw.csv > profiles/NAME.md
"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("smsp__inst_executed.sum", "inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    cols = [(hdr.index(m), short) for m, short in WANT if m in hdr]
    print("| kernel | " + " | ".join(f"{s} [{units[i]}]" if units[i] else s for i, s in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        name = r[kn].split("(")[0].replace("void ", "").replace("plaid::", "")[:48]
        vals = []
        for i, _ in cols:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
