#!/usr/bin/env python
"""Top-N SASS lines by warp-stall samples from an .ncu-rep source page, with the dominant stall reasons."""
import csv, subprocess, sys, io
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--print-source", sys.argv[3]] if len(sys.argv) > 3 else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si = h.index("# Samples")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
body = [r for r in rows[2:] if len(r) == len(h)]
f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
tot = sum(f(r[si]) for r in body)
order = sorted(range(len(body)), key=lambda i: -f(body[i][si]))
print("total samples", tot)
for i in order[:n]:
    r = body[i]
    st = sorted(((f(r[c]), h[c][6:]) for c in stall_cols), reverse=True)[:3]
    print("%5.1f%% #%-5d %-70s %s" % (100 * f(r[si]) / tot, i, r[1].strip()[:70], " ".join("%s=%d" % (b, a) for a, b in st if a)))
