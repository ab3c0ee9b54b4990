#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (developer tool)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1
    tot[name][1] += v
T = sum(v for _, v in tot.values())
print(f"total {T:.1f} us over {sum(n for n, _ in tot.values())} launches")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={n:4d}  {k}")
