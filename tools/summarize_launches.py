#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: tools/summarize_launches.py gpurun_out/launches_X.csv [> profiles/launches_X.summary.txt]"""
import csv, re, sys, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "")
    ns = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit") == "us": ns *= 1e3
    if r.get("Metric Unit") == "ms": ns *= 1e6
    rows.append((name, ns))
tot = sum(n for _, n in rows)
agg = collections.OrderedDict()
for k, n in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += n
print("launches %d   total %.3f ms (cold-cache, serialised under ncu: compare shares)" % (len(rows), tot / 1e6))
print("%-44s %6s %10s %7s" % ("kernel", "count", "ms", "share"))
for k, (c, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %6d %10.3f %6.1f%%" % (k[:44], c, n / 1e6, 100 * n / tot))
