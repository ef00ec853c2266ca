#!/usr/bin/env python
"""Hot instructions of a kernel from an ncu source page.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:render_bwd > src.csv
    python tools/ncu_source_hot.py src.csv [min share of executed instructions, default 0.002]

Prints the total warp-level instruction and sample counts, the stall reasons summed over the kernel, and every SASS
instruction whose executed count exceeds the share: index, SASS, executed count (thousands), samples, average active
threads, its two largest stall reasons.  (The CSV lists every instruction of a kernel twice when the report holds two
launches of it: the totals printed are then twice the per-launch figure.)
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
tot = sum(int(r[idx["Instructions Executed"]]) for r in data)
tots = sum(int(r[idx["# Samples"]]) for r in data)
print("instr", tot, "samples", tots, "n", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:10])
for i, r in enumerate(data):
    e = int(r[idx["Instructions Executed"]])
    s = int(r[idx["# Samples"]])
    if e > tot * thr:
        top = sorted(((int(r[idx[x]]), x) for x in stalls), reverse=True)[:2]
        print(i, r[idx["Source"]].strip()[:58].ljust(58), e // 1000, s, r[idx["Avg. Threads Executed"]],
              [(b[6:], a) for a, b in top if a > 0])
