"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  python tools/launch_summary.py list.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v, u = float(r[iv].replace(",", "")), r[iu]
    us = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v * 1e6 if u in ("s", "second") else v
    agg[r[ik]].append(us)
tot = sum(sum(v) for v in agg.values())
print(f"# total kernel time {tot / 1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v) / 1e3:10.3f} ms {100 * sum(v) / tot:5.1f}% {len(v):5d} launches  avg {sum(v) / len(v):9.1f} us  {k[:110]}")
