#!/usr/bin/env python
"""Opcode mix, stall samples and shared-memory wavefronts per SASS opcode from an .ncu-rep captured with --import-source on.

usage: python tools/ncu_opmix.py gpurun_out/x.ncu-rep [units]     (units: divide the counts, e.g. the number of warp walks)
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[1], rows[2:]
    iS, iE, iSm, iW = (hdr.index(k) for k in ("Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared"))
    tot, samp, wf = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS].strip())
        op = m.group(2) if m else r[iS].strip()
        tot[op.split('.')[0]] += int(r[iE])
        samp[op.split('.')[0]] += int(r[iSm])
        wf[op] += int(r[iW])
    T, S = sum(tot.values()), sum(samp.values())
    print("instructions executed %d (%.0f per unit), %d static SASS lines" % (T, T / units, len(data)))
    for k, v in tot.most_common(24):
        print("%-12s %12.0f /unit %5.1f%%  stall samples %5.1f%%" % (k, v / units, 100.0 * v / T, 100.0 * samp[k] / S))
    print("shared wavefronts /unit:", {k: round(v / units, 1) for k, v in wf.items() if v})


if __name__ == "__main__":
    main()
