#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report (one launch per row) as a markdown table.
  python scripts/summarize_full.py gpurun_out/prof_img_fwd.ncu-rep profiles/r1h_image_ncu_full_fwd.md "title" """
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
cols = [(w, hdr.index(w)) for w in want if w in hdr]
with open(out, "w") as f:
    f.write("# %s\n\n`ncu --set full --clock-control none` (%s), one launch per row.\n\n" % (title, rep.split("/")[-1]))
    f.write("| " + " | ".join(w.split(".avg")[0].split(".sum")[0] for w, _ in cols) + " |\n|" + "---|" * len(cols) + "\n")
    for row in r[2:]:
        f.write("| " + " | ".join((row[i].split("(")[0][-40:] if w == "Kernel Name" else row[i] + " " + units[i]) for w, i in cols) + " |\n")
print(open(out).read()[:6000])
