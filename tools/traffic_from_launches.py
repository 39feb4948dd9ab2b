"""DRAM traffic per evaluation of each kernel family, from an ncu --csv launch list that carries
dram__bytes_read.sum / dram__bytes_write.sum per launch (bench.py --profile = 2 evaluations).
usage: python tools/traffic_from_launches.py launches.csv workload n_evals out.json"""
import collections
import csv
import json
import sys

src, workload, nev, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(src)))
hdr = None
acc = collections.defaultdict(lambda: {"read": 0.0, "write": 0.0, "launches": 0, "ns": 0.0})
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "").strip()
        v = float(d["Metric Value"].replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(d["Metric Unit"], 1.0)
        if d["Metric Name"] == "dram__bytes_read.sum":
            acc[name]["read"] += v * scale
        elif d["Metric Name"] == "dram__bytes_write.sum":
            acc[name]["write"] += v * scale
        elif d["Metric Name"] == "gpu__time_duration.sum":
            acc[name]["ns"] += v * scale
            acc[name]["launches"] += 1
res = {}
try:
    res = json.load(open(out))
except (OSError, ValueError):
    pass
res[workload] = {k: {"dram_bytes_per_eval": (v["read"] + v["write"]) / nev, "launches_per_eval": v["launches"] / nev,
                     "ncu_ms_per_eval": v["ns"] / nev / 1e6} for k, v in acc.items()}
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(res[workload], indent=1))
