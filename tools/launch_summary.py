#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the file kept under profiles/)."""
import collections, csv, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) > vi:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n[:60]:60s} n={len(v):3d} mean={sum(v) / len(v) / 1000:9.1f} us  last={v[-1] / 1000:9.1f} min={min(v) / 1000:9.1f}")
