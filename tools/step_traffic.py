#!/usr/bin/env python
"""profiles/<round>_step_traffic.json from an ncu pass over one training step
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`, tools/gpu_suite.sh traffic):
per kernel, launches per step, average duration and average DRAM bytes per launch.  bench.py reads the file to fill
`roofline.traffic` for the dominant kernel class (the capture is per launch, cold-cache, serialised)."""
import collections
import csv
import json
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_traffic.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r1_step_traffic.json"
lines = [l for l in open(src) if not l.startswith("==")]
by = collections.OrderedDict()
for r in csv.DictReader(lines):
    k = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip())
    by.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
agg = collections.OrderedDict()
for (_, name), m in by.items():
    d = m["gpu__time_duration.sum"]
    us = d[0] / 1e3 if d[1] == "ns" else (d[0] * 1e3 if d[1] == "ms" else d[0])

    def nbytes(x):
        v, u = m.get(x, (0.0, "byte"))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    a = agg.setdefault(name, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    a["launches"] += 1
    a["us"] += us
    a["dram_read"] += nbytes("dram__bytes_read.sum")
    a["dram_write"] += nbytes("dram__bytes_write.sum")
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 "one k2 training step (tools/profile_step.py); per-launch averages, cold-cache and serialised",
       "kernels": {}}
for name, a in agg.items():
    n = a["launches"]
    out["kernels"][name] = {"launches_per_step": n, "avg_us": a["us"] / n,
                            "avg_dram_bytes": (a["dram_read"] + a["dram_write"]) / n,
                            "avg_dram_read_bytes": a["dram_read"] / n, "avg_dram_write_bytes": a["dram_write"] / n,
                            "dram_GBps": (a["dram_read"] + a["dram_write"]) / a["us"] / 1e3}
json.dump(out, open(dst, "w"), indent=1)
print(f"wrote {dst}: {len(out['kernels'])} kernels")
