#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
launches, total device time and share (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections
import csv
import sys


def main(path, only_plaid=False):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")[:70]
        tot[name][0] += 1
        tot[name][1] += float(r[-1].replace(",", ""))
    total = sum(v[1] for v in tot.values())
    print("| kernel | launches | total [us] | share |")
    print("|---|---|---|---|")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if only_plaid and "plaid" not in name and "_kernel" not in name:
            continue
        print(f"| {name} | {n} | {t / 1e3:.1f} | {t / total:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
