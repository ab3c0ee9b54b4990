#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log:
launch count, average duration, average DRAM traffic and the resulting GB/s (developer tool)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else "cs::"
lines = [l for l in open(path) if not l.startswith("==")]
by = collections.OrderedDict()
for r in csv.DictReader(lines):
    k = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"])[:60])
    by.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
agg = collections.OrderedDict()
for (_, name), m in by.items():
    if only not in name:
        continue
    d = m["gpu__time_duration.sum"]
    t = d[0] / 1e3 if d[1] == "ns" else (d[0] * 1e3 if d[1] == "ms" else d[0])

    def nbytes(x):
        if x not in m:
            return 0.0
        v, u = m[x]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] += nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
for k, (n, t, b) in agg.items():
    print(f"{k:62s} n={n:3d} avg {t / n:8.1f} us  dram {b / n / 1e6:8.1f} MB  -> {b / n / (t / n) / 1e3:7.1f} GB/s")
