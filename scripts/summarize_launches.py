#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms total (ncu per-launch durations: cold-cache, serialised)")
print(f"{'kernel':90s} {'n':>5s} {'ms':>10s} {'share':>7s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:90s} {n:5d} {ms:10.3f} {ms / tot:7.3f}")
