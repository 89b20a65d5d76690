"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time per kernel name over the LAST `n` launches
(one step when n = launches per step).  Usage: python tools/launch_summary.py launches.csv [n_last]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows)
rows = rows[-n:]
agg = OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("sky::", "")[:70]
    t = float(r[14])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:9.1f} us {100 * t / tot:5.1f}%  x{c:<4d} {name}")
