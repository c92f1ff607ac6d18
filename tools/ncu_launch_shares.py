"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the captured window).
    python tools/ncu_launch_shares.py gpurun_out/launches_<tag>.csv > profiles/r01_launch_shares_<tag>.txt"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(rows)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
tot = 0.0
n = 0
for row in r:
    name = re.sub(r"\(.*", "", row[ki])
    us = float(row[vi].replace(",", "")) / 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
    n += 1
print(f"{n} launches, {tot / 1e3:.2f} ms  (cold-cache, serialised under the profiler: compare shares, not absolutes)\n")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {c:5d} launches {100 * us / tot:5.1f}%  {name}")
