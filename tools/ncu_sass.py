#!/usr/bin/env python
"""SASS-level stall profile of one kernel from an ncu report (--set full --import-source on).
    python tools/ncu_sass.py <report.ncu-rep> <kernel regex> [min share %]
Prints the stall-reason totals and every instruction holding more than the given share of the
warp-stall samples, with its dominant stall reasons and the CUDA source line it maps to."""
import collections, csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "Line No")]
h = rows[hi[0]]
end = hi[1] if len(hi) > 1 else len(rows)
def I(x):
    try: return int(x)
    except ValueError: return 0
ai, src = h.index("Address"), h.index("Source")
si, ii = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
stall = [j for j, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
sass = [r for r in rows[hi[0] + 1:end] if len(r) > ii and r[ai]]
tot = sum(I(r[si]) for r in sass) or 1
toti = sum(I(r[ii]) for r in sass) or 1
cat = collections.Counter()
for r in sass:
    for j in stall: cat[h[j][6:]] += I(r[j])
print(len(sass), "instructions,", tot, "samples,", toti, "warp instructions executed")
print("stall reasons:", [(k, round(100 * v / tot, 1)) for k, v in cat.most_common(9)])
ops = collections.Counter()
for r in sass:
    op = r[src].split()
    op = [o for o in op if not o.startswith("@")]
    if op: ops[op[0].split(".")[0]] += I(r[ii])
print("executed mix:", [(k, round(100 * v / toti, 1)) for k, v in ops.most_common(14)])
for k, r in enumerate(sass):
    s = I(r[si])
    if s > tot * minshare / 100:
        why = {h[j][6:]: I(r[j]) for j in stall if I(r[j]) > s * 0.2}
        print(f"{k:5d} {r[src][:64]:64s} {100 * s / tot:5.1f}%  x{I(r[ii])}  {why}")
