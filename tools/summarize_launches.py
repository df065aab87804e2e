#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv, sys, re, collections
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
agg = collections.OrderedDict()
for name, ns, grid, blk in rows:
    short = re.sub(r"\(.*$", "", name)
    short = re.sub(r"<.*", "", short).replace("void ", "")
    d = agg.setdefault(short, [0.0, 0])
    d[0] += ns; d[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"# {path}: {len(rows)} launches, {tot/1e6/steps:.3f} ms per step (serialised, cold-cache), {steps} step(s)")
print(f"{'kernel':70s} {'launches/step':>13s} {'ms/step':>9s} {'share':>7s}")
for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{k[:70]:70s} {n/steps:13.1f} {ns/1e6/steps:9.3f} {100*ns/tot:6.1f}%")
