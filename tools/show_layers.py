#!/usr/bin/env python
"""Pretty-print a bench.py --per-layer table."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
tot = [0.0] * 7
print("%2s %-40s" % ("#", "scope"), " ".join("%8s" % h for h in rows[0][2:]))
for r in rows[1:]:
    print("%2s %-40s" % (r[0], r[1]), " ".join("%8.3f" % float(x) for x in r[2:]))
    for i, x in enumerate(r[2:]): tot[i] += float(x)
print("%-43s" % "total", " ".join("%8.3f" % t for t in tot), " sum %.3f" % sum(tot))
