#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name + grid size.
  python scripts/summarize_launches.py gpurun_out/image_launches.csv [out.md]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
tot, n = 0.0, 0
for r in csv.DictReader(lines):
    try:
        us = float(r["Metric Value"].replace(",", "")) / 1e3
    except Exception:
        continue
    m = re.search(r"(k_\w+(<[^(]*>)?)", r["Kernel Name"])
    name = m.group(1) if m else r["Kernel Name"].split("(")[0][-60:]
    key = (name, r.get("Grid Size", "?"))
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += us
    tot += us
    n += 1
out = ["%d launches, %.1f us in total (cold-cache, serialised: compare SHARES)\n" % (n, tot),
       "| kernel | grid | launches | us total | us each | share |", "|---|---|---:|---:|---:|---:|"]
for (name, grid), (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %s | %d | %.1f | %.1f | %.1f%% |" % (name, grid.replace(" ", ""), c, v, v / c, 100 * v / tot))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
