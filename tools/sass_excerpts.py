"""SASS evidence for profiles/: per kernel of libbppgpu.so, how often the mnemonics that prove the data path occur
(DMMA = FP64 tensor cores, UBLKCP = cp.async.bulk / TMA engine, SYNCS = mbarrier, LDGSTS = cp.async, ACQBULK / griddepcontrol
= programmatic dependent launch), plus the first occurrence of each with its neighbours.  Usage:
    python tools/sass_excerpts.py bpp_phyl_b200/lib/libbppgpu.so > profiles/r2_sass_excerpts.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["DMMA", "DFMA", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDS", "LDG", "STG", "BAR", "ACQBULK", "PREEXIT", "SHFL"]
so = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, body = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        body[cur] = []
    elif cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        body[cur].append(line.strip())
print(f"# SASS mnemonic counts per kernel of {so} (cuobjdump -sass, sm_100a)\n")
for fn, lines in body.items():
    ops = collections.Counter()
    first = {}
    for i, l in enumerate(lines):
        m = re.match(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                ops[k] += 1
                first.setdefault(k, i)
    name = demangle(fn)
    name = name[:150] + ("..." if len(name) > 150 else "")
    print(f"## {name}\n   {len(lines)} instructions; " + ", ".join(f"{k} {ops[k]}" for k in KEYS if ops[k]))
    for k in ("DMMA", "UBLKCP", "UTMALDG", "SYNCS", "ACQBULK", "PREEXIT"):
        if k in first:
            i = first[k]
            for l in lines[max(0, i - 1):i + 2]:
                print("      " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l))
    print()
