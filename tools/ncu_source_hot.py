"""Hottest source lines (stall samples) of one kernel instance in an .ncu-rep captured with --import-source on.
usage: python tools/ncu_source_hot.py rep.ncu-rep '<substring of kernel name>' [instance] [top]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# sections: (File Path, Function Name) pairs; a new kernel instance starts when common file order restarts
sections = []
i = 0
while i < len(rows):
    r = rows[i]
    if len(r) == 2 and r[0] == "File Path" and i + 1 < len(rows) and rows[i + 1][0] == "Function Name":
        sections.append({"file": r[1].split("/")[-1], "func": rows[i + 1][1], "rows": []})
        i += 2
        continue
    if sections:
        sections[-1]["rows"].append(r)
    i += 1
inst, seen = [], set()
for s in sections:
    key = (s["func"], s["file"])
    if not inst or key in seen or s["func"] != inst[-1][0]["func"]:
        inst.append([])
        seen = set()
    seen.add(key)
    inst[-1].append(s)
inst = [x for x in inst if pat in x[0]["func"]]
sel = inst[which]
out = []
for s in sel:
    hdr = None
    for r in s["rows"]:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            d = dict(zip(hdr, r))
            try:
                st = {k[6:]: int(d[k]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and d[k] not in ("", "0")}
                out.append((int(d["# Samples"]), int(d["Instructions Executed"]), s["file"], d["Line No"], r[1].strip()[:90], st))
            except ValueError:
                pass
out.sort(key=lambda o: -o[0])
print(sel[0]["func"], "samples", sum(o[0] for o in out), "warp-instr", sum(o[1] for o in out))
for o in out[:top]:
    print(o)
