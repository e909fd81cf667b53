"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file F` launch list: per kernel launches, total, share.
Usage: python tools/ncu_launch_summary.py F [last_n_launches]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    out.append((r[idx["Kernel Name"]], float(r[idx["Metric Value"]].replace(",", "")), r[idx["Grid Size"]]))
if len(sys.argv) > 2:
    out = out[-int(sys.argv[2]):]
agg = OrderedDict()
for k, v, g in out:
    a = agg.setdefault(k[:92], [0, 0.0, g])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':94s}{'launches':>9s}{'total ns':>12s}{'ns/launch':>11s}{'share':>7s}  grid")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:94s}{a[0]:9d}{a[1]:12.0f}{a[1] / a[0]:11.0f}{a[1] / tot:7.3f}  {a[2]}")
print(f"{'total':94s}{'':9s}{tot:12.0f}")
