"""Per-kernel breakdown of ONE step from an ncu launch list of bench.py (launches between the last two L2 flushes).
Usage: python tools/step_breakdown.py launches.csv [top_n]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
idx = [i for i, r in enumerate(rows) if "FillFunctor<unsigned char>" in r[4]]
a, b = idx[-2], idx[-1]
step = rows[a + 1:b]
tot = sum(float(r[14]) for r in step) / 1e3
print(f"{len(step)} launches between the last two L2 flushes, {tot:.1f} us (cold-cache, serialised)")
agg = {}
for r in step:
    name = r[4].split("(")[0].replace("void ", "").replace("sky::", "")[:48]
    key = (name, r[8])
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += float(r[14]) / 1e3
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1]:8.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:<3d} {k[0]:48s} grid {k[1]}")
