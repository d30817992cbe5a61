#!/usr/bin/env python
"""Turn the ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag>      # e.g. r01a
reads gpurun_out/launches.csv (gpu__time_duration launch list) and gpurun_out/prof.ncu-rep (--set full)
writes profiles/<tag>_launches.md and profiles/<tag>_ncu_full.md
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki].split("(")[0].replace("void ", "")[:80], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, tag + "_launches.md"), "w") as f:
    f.write("# %s: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py "
            "--steps 2 --warmup 1 --no-cpu-baseline` (cold-cache, serialised: compare SHARES, not absolutes).\n\n")
    f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    f.write("\ntotal %.1f us over %d launches\n" % (tot, sum(v[0] for v in agg.values())))

rep = os.path.join(G, "prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    kn = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, tag + "_ncu_full.md"), "w") as f:
        f.write("# %s: ncu --set full --clock-control none (one launch per row; values per launch)\n\n" % tag)
        f.write("| kernel | " + " | ".join(w for w, _ in idx) + " |\n|---|" + "---:|" * len(idx) + "\n")
        for r in rr[2:]:
            f.write("| `%s` | " % r[kn].split("(")[0].replace("void ", "")[:60]
                    + " | ".join("%s %s" % (r[i][:12], units[i]) for _, i in idx) + " |\n")
print("wrote profiles/%s_*" % tag)
