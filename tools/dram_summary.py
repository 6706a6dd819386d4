#!/usr/bin/env python
"""Per-kernel DRAM traffic and time of ONE frame from an ncu launch list taken with

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...

Usage: dram_summary.py launches.csv [n_frames] [out.json]
`n_frames` = frames covered by the list (launch counts are divided by it).  Writes the per-frame totals bench.py reports as
`roofline.traffic` (MLP kernel, per launch) and `roofline.traffic_all_kernels_per_frame` into out.json."""
import collections
import csv
import json
import sys

path = sys.argv[1]
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    if r[mi] == "gpu__time_duration.sum":
        a["ns"] += v
        a["n"] += 1
    elif r[mi] == "dram__bytes_read.sum":
        a["rd"] += v
    elif r[mi] == "dram__bytes_write.sum":
        a["wr"] += v
tot = {"ns": 0.0, "rd": 0.0, "wr": 0.0}
print(f"{'kernel':52s} {'launches/frame':>14s} {'us/frame':>10s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{n[:52]:52s} {a['n'] / n_frames:14.1f} {a['ns'] / n_frames / 1e3:10.1f} {a['rd'] / n_frames / 1e6:11.2f} {a['wr'] / n_frames / 1e6:11.2f}")
    for k in tot:
        tot[k] += a[k] / n_frames
print(f"{'ALL KERNELS, per frame':52s} {'':14s} {tot['ns'] / 1e3:10.1f} {tot['rd'] / 1e6:11.2f} {tot['wr'] / 1e6:11.2f}")
if len(sys.argv) > 3:
    mlp = next((a for n, a in agg.items() if "mlp_tc" in n), None)
    out = {
        "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ({path}), {n_frames} frame(s)",
        "dram_bytes_per_launch": None if mlp is None else (mlp["rd"] + mlp["wr"]) / max(mlp["n"], 1),
        # the path's own kernels only: bench.py's L2 flush (a 256 MB fill) and its clock probe are part of the list but not of the path
        "dram_bytes_per_frame": sum((a["rd"] + a["wr"]) / n_frames for n, a in agg.items() if n.startswith("dsn::")),
        "per_kernel_per_frame": {n: {"launches": a["n"] / n_frames, "us": a["ns"] / n_frames / 1e3, "dram_bytes": (a["rd"] + a["wr"]) / n_frames}
                                 for n, a in agg.items()},
    }
    json.dump(out, open(sys.argv[3], "w"), indent=1)
