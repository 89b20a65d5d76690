"""Summarise an ncu source-page CSV: top SASS instructions by stall samples, with their dominant stall reasons.
Usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1], " total samples:", tot)
agg = {}
for s in stall_cols:
    agg[s] = sum(int(r[ix[s]] or 0) for r in body)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[s]] or 0), s) for s in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} exec={r[ix['Instructions Executed']]:>9} "
          f"confl={r[ix['L1 Wavefronts Shared Excessive']]:>9} {r[ix['Source']].strip()[:70]:70s} "
          + " ".join(f"{s[6:]}={v}" for v, s in st if v))
