#!/usr/bin/env python
"""Print the last N launches of an `ncu --metrics gpu__time_duration.sum --csv` launch list in order (one training step)."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = 0.0
for k, us, g, b in rows[-n:]:
    tot += us
    print("%8.1f us  %-28s %-14s %s" % (us, g, b, k[:110]))
print("sum %.1f us over %d launches" % (tot, min(n, len(rows))))
