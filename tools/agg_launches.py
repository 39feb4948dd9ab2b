"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

import numpy as np

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(list)
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            agg[d["Kernel Name"][:70]].append(float(d["Metric Value"]))
        except ValueError:
            pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    pc = np.percentile(v, [5, 50, 95]) / 1e3
    print(f"{k:70s} n={len(v):5d} sum={sum(v)/1e3:10.1f}us mean={sum(v)/len(v)/1e3:8.1f}us p5/p50/p95={pc[0]:.1f}/{pc[1]:.1f}/{pc[2]:.1f}")
