#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep by stall reason and by opcode."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si = h.index("# Samples")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
body = [r for r in rows[2:] if len(r) == len(h)]
f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
tot = sum(f(r[si]) for r in body)
by_reason = collections.Counter(); by_op = collections.Counter()
for r in body:
    for c in stall_cols:
        by_reason[h[c]] += f(r[c])
    op = r[1].strip().split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    by_op[op.split(".")[0]] += f(r[si])
print("samples", tot, "instructions", len(body))
print("by reason:", ", ".join("%s=%.1f%%" % (k[6:], 100 * v / tot) for k, v in by_reason.most_common(10)))
print("by opcode:", ", ".join("%s=%.1f%%" % (k, 100 * v / tot) for k, v in by_op.most_common(16)))
