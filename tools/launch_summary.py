"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count and mean duration.
    python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ki][:60], r[gi]), []).append(float(r[vi].replace(",", "")))
total = sum(sum(v) for v in agg.values())
for (k, g), v in agg.items():
    print(f"{k:60s} grid {g:>16s}  x{len(v):4d}  mean {sum(v) / len(v) / 1000:9.2f} us  share {100 * sum(v) / total:5.1f} %")
