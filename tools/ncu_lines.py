#!/usr/bin/env python
"""Per-source-line stall/instruction shares of one kernel from an ncu report.
    python tools/ncu_lines.py <report.ncu-rep> <source file> [top]"""
import collections, csv, subprocess, sys
rep, srcfile = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
li, si, ii = h.index("Line No"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
ai = h.index("Address")
agg = collections.defaultdict(lambda: [0, 0])
tot = toti = 0
for r in rows[hi + 1:]:
    if len(r) <= ii or not r[ai]:      # SASS rows only (they carry an address)
        continue
    try:
        s, n = int(r[si]), int(r[ii])
    except ValueError:
        continue
    agg[r[li]][0] += s; agg[r[li]][1] += n
    tot += s; toti += n
lines = open(srcfile).read().split("\n")
print("total samples", tot, "warp instructions", toti)
for ln, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    try:
        txt = lines[int(ln) - 1].strip()[:100]
    except Exception:
        txt = ""
    print(f"{ln:>5s} stall {100 * s / max(tot, 1):5.1f}%  inst {100 * n / max(toti, 1):5.1f}%  {txt}")
