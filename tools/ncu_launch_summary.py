"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`):
launches, total and average duration and share per kernel.  Usage: python tools/ncu_launch_summary.py x.csv"""
import csv
import collections
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    a = agg.setdefault(name, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for k, a in agg.items() if "synth" not in k)
print(f"{'kernel':34s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}  grid/block")
for k, a in agg.items():
    share = "   -  " if "synth" in k else f"{100 * a[1] / tot:5.1f}%"
    print(f"{k:34s} {a[0]:8d} {a[1] / 1e3:10.3f} {a[1] / a[0]:10.1f} {share:>7s}  {a[2]} {a[3]}")
