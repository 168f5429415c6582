#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches,
total device time and share.  Usage: python tools/summarize_launches.py launches.csv [--from-first dd::] [> summary.md]
(--from-first PREFIX drops everything before the first kernel whose name starts with PREFIX, e.g. weight initialisation)"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        rows.append((name, ns))
if "--from-first" in sys.argv:
    pre = sys.argv[sys.argv.index("--from-first") + 1]
    first = next((i for i, (n, _) in enumerate(rows) if n.startswith(pre)), 0)
    rows = rows[first:]
tot = sum(ns for _, ns in rows)
agg = collections.OrderedDict()
for name, ns in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
print(f"| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {ns / 1e3:.1f} | {ns / n / 1e3:.2f} | {ns / tot:.3f} |")
print(f"| **all** | {len(rows)} | {tot / 1e3:.1f} | | 1.000 |")
