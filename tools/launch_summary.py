"""Per-kernel digest of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[0].isdigit():
        continue
    agg.setdefault(r[h.index("Kernel Name")], []).append(float(r[-1].replace(",", "")))
tot = sum(v[-1] for v in agg.values())
for k, v in agg.items():
    print(f"{k[:70]:70s} n={len(v):3d} last={v[-1]/1e3:9.1f} us  {100*v[-1]/tot:5.1f}%")
print(f"one launch of each: {tot/1e3:.1f} us")
