#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into tracked files under profiles/.
  python scripts/summarize_ncu.py <tag>     # e.g. r1b
reads gpurun_out/launches.csv (gpu__time_duration per launch) and gpurun_out/prof_<round>.ncu-rep (--set full)."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rep = sys.argv[2] if len(sys.argv) > 2 else "prof_r1.ncu-rep"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -> per-step table
rows = []
with open(os.path.join(ROOT, "gpurun_out", "launches.csv")) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    try:
        rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", "")) / 1e3))
    except Exception:
        pass
idx = [i for i, r in enumerate(rows) if "k_clip_sgd" in r[0]]
step = rows[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else rows
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for n, v in step:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
with open(os.path.join(out, tag + "_launches_step.md"), "w") as f:
    f.write("# %s — ncu launch list of ONE fused inner step (bench.py --steps 2 --warmup 1 --no-cpu --no-e2e)\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: "
            "compare SHARES.  %d launches, sum %.1f us.\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n" % (len(step), tot))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (n.replace("lagvae::", "").replace("<unnamed>::", ""), c, v, 100 * v / tot))
    f.write("\n## launch order\n\n```\n")
    for n, v in step:
        f.write("%9.1f  %s\n" % (v, n.replace("lagvae::", "").replace("<unnamed>::", "")))
    f.write("```\n")

# ---- full capture -> key metrics
path = os.path.join(ROOT, "gpurun_out", rep)
if os.path.exists(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
    cols = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(out, tag + "_ncu_full.md"), "w") as f:
        f.write("# %s — `ncu --set full --clock-control none` (one launch per row; %s)\n\n" % (tag, rep))
        f.write("| " + " | ".join(w.split(".")[0] for w, _ in cols) + " |\n|" + "---|" * len(cols) + "\n")
        for row in r[2:]:
            f.write("| " + " | ".join((row[i].split("(")[0][-48:] if w == "Kernel Name" else row[i] + " " + units[i]) for w, i in cols) + " |\n")
print("wrote", [x for x in os.listdir(out) if x.startswith(tag)])
