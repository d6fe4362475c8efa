"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches, total and mean
device time, share of the whole list.  usage: python tools/ncu_launch_table.py launches.csv [skip_first_n]"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^.*unnamed>::", "", name)
        rows.append((name[:70], v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = sum(v for _, v in rows) or 1.0
agg = collections.OrderedDict()
for n, v in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"{'kernel':70s} {'n':>5s} {'total us':>10s} {'mean us':>9s} {'share':>6s}")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:70s} {c:5d} {v:10.1f} {v / c:9.2f} {100 * v / tot:5.1f}%")
print(f"{'TOTAL':70s} {len(rows):5d} {tot:10.1f}")
