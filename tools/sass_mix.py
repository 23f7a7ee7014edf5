#!/usr/bin/env python
"""Dynamic instruction mix of one kernel from an ncu report's source page:
   ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; sass_mix.py src.csv [units]
Prints executed warp instructions per opcode, per unit (default 1,000,000 units = one bench launch of pair hashes)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
hdr = rows[1]
src, ex, samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, stall = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= ex:
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[src])
    if not m:
        continue
    op = m.group(1)
    key = op if op.startswith(("IMAD", "DFMA", "DADD", "DMUL")) else op.split(".")[0]
    mix[key] += int(r[ex])
    stall[key] += int(r[samp])
tot, stot = sum(mix.values()), sum(stall.values()) or 1
print(f"# executed warp instructions x 32 / {units:g} units   (total {tot * 32 / units:,.0f} per unit)")
for k, v in mix.most_common(30):
    print(f"{k:22s} {v * 32 / units:10.1f} per unit {100 * v / tot:6.2f}%   stall samples {100 * stall[k] / stot:5.1f}%")
