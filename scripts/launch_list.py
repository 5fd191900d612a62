"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / median / total.
usage: python scripts/launch_list.py gpurun_out/launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print(f"{k:70s} n={len(v):4d} med={sorted(v)[len(v) // 2]:10.1f} sum={sum(v):12.1f} {rows[1][ui]}")
