#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, share).
usage: launch_list_summary.py launches.csv out.txt "command that was profiled" """
import collections
import csv
import sys

src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[mu], 1.0)
    a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
with open(out, "w") as f:
    f.write(f"# per-kernel totals of {src}\n# command: {cmd}\n# (ncu serialises launches and runs them cold-cache: shares are meaningful, absolutes are not bench values)\n")
    f.write(f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'ms/launch':>10s}\n")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k[:62]:62s} {c:8d} {t:10.3f} {100 * t / tot:6.1f}% {t / c:10.4f}\n")
print(open(out).read())
