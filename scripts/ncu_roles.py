#!/usr/bin/env python3
"""Per-address stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv`: prints the
hottest instructions and, given address ranges, the samples / stall mix per warp role.
usage: ncu_roles.py src.csv [name:lo:hi ...]   (hex offsets from the kernel's first instruction)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
cols = {k: hdr.index(k) for k in names}
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), r[isrc], int(r[isamp] or 0), int(r[iex] or 0), {k: int(r[c] or 0) for k, c in cols.items()}))
    except Exception:
        pass
base = data[0][0]
tot = sum(d[2] for d in data)
print('total samples', tot, 'instructions', sum(d[3] for d in data))
for d in sorted(data, key=lambda d: -d[2])[:25]:
    st = {k[6:]: v for k, v in d[4].items() if v > 0.15 * d[2]}
    print('%5x %6d %5.1f%% ex=%9d %-58s %s' % (d[0] - base, d[2], 100 * d[2] / tot, d[3], d[1][:58], st))
for spec in sys.argv[2:]:
    name, lo, hi = spec.split(':')
    lo, hi = int(lo, 16), int(hi, 16)
    sel = [d for d in data if lo <= d[0] - base < hi]
    agg = {}
    for d in sel:
        for k, v in d[4].items():
            agg[k[6:]] = agg.get(k[6:], 0) + v
    print(name, 'samples', sum(d[2] for d in sel), 'instr', sum(d[3] for d in sel), sorted(agg.items(), key=lambda kv: -kv[1])[:6])
