#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep + launch list CSV into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> (e.g. r1b) <round label> (e.g. r01)"""
import collections
import csv
import io
import os
import subprocess
import sys

tag, label = sys.argv[1], sys.argv[2]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

# 1) launch list: per-kernel totals and shares (cold-cache, serialised: shares matter, not absolutes)
lines = [l for l in open(os.path.join(G, "launches_%s.csv" % tag)) if l.startswith('"')]
agg = collections.OrderedDict()
raw_rows = []
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    raw_rows.append((row["ID"], name, "%.3f" % v))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(OUT, "%s_launch_list.csv" % label), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 python bench.py --steps 1 --warmup 1 --batch %d --profile-mode\n" % batch)
    f.write("id,kernel,duration_us\n")
    for r in raw_rows:
        f.write(",".join(r) + "\n")
with open(os.path.join(OUT, "%s_launch_shares.md" % label), "w") as f:
    f.write("# Kernel shares of the profiled window (ncu launch list, 400 launches of `bench.py --batch %d`)" % batch + "\n\n")
    f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))

# 2) per-kernel metric summaries from the --set full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
summary = {}
with open(os.path.join(OUT, "%s_kernel_metrics.md" % label), "w") as f:
    f.write("# ncu --set full --clock-control none captures (B = %d measurements = %d frames of 128x128 per launch)\n\n" % (batch, 8 * batch))
    for kind in ["hidden", "last", "first", "prep", "gram", "solve", "mix"]:
        rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (kind, tag))
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        f.write("## %s\n\n| metric | unit | value(s) |\n|---|---|---|\n" % kind)
        for i, h in enumerate(hdr):
            if h == "Kernel Name":
                f.write("| kernel | | `%s` |\n" % rows[2][i][:90])
            if h in WANT:
                f.write("| %s | %s | %s |\n" % (h, units[i], ", ".join(r[i] for r in rows[2:])))
        f.write("\n")
        # machine-readable: DRAM bytes per launch (bench.py reads roofline.traffic from here)
        def col(name):
            i = hdr.index(name)
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "%": 1}.get(units[i], 1)
            vals = [float(r[i].replace(",", "")) * scale for r in rows[2:]]
            return sum(vals) / len(vals)
        summary[kind] = {"batch": batch, "kernel": rows[2][hdr.index("Kernel Name")][:80],
                         "dram_bytes_per_launch": col("dram__bytes_read.sum") + col("dram__bytes_write.sum"),
                         "dram_read_bytes": col("dram__bytes_read.sum"), "dram_write_bytes": col("dram__bytes_write.sum"),
                         "duration_us_under_ncu": col("gpu__time_duration.sum"),
                         "dram_pct_of_peak": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         "tensor_pipe_pct": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
import json
json.dump(summary, open(os.path.join(OUT, "%s_kernel_metrics.json" % label), "w"), indent=1)
print("wrote", os.listdir(OUT))
