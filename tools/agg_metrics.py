"""Per-kernel summary of an ncu --csv log holding several metrics per launch (time, instructions, DMMA pipe %)."""
import collections
import csv
import sys

import numpy as np

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = collections.defaultdict(dict)
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        per[d["ID"]][d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
        per[d["ID"]]["k"] = d["Kernel Name"][:60]
by = collections.defaultdict(list)
for v in per.values():
    by[v["k"]].append(v)
for k, vs in by.items():
    print(k, "n=%d" % len(vs))
    for m in vs[0]:
        if m == "k":
            continue
        a = np.array([v[m] for v in vs])
        print("   %-80s mean %.4g  p5/50/95 %s" % (m, a.mean(), np.percentile(a, [5, 50, 95]).round(2)))
