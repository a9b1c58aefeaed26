#!/usr/bin/env python
"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` output. usage: ncu_hot.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        samples = int(r[col["# Samples"]])
    except ValueError:
        continue
    data.append((idx, samples, r))
tot = sum(d[1] for d in data)
print(f"total samples {tot}, instructions {len(data)}")
for idx, samples, r in sorted(data, key=lambda d: -d[1])[:top]:
    st = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:3]
    print(f"{idx:5d} {samples:6d} {100.0 * samples / tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} exec={r[col['Instructions Executed']]:>9s} " + " ".join(f"{n}:{v}" for v, n in st if v))
