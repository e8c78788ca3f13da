#!/usr/bin/env python
"""Top stall sites (SASS lines by warp-stall samples) of each kernel in an .ncu-rep captured with --import-source on:
   python tools/ncu_stalls.py rep [kernel regex] [top n]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', raw)
for blk in blocks[1:]:
    lines = blk.split("\n")
    name = lines[0][:110]
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    if len(rows) < 2:
        continue
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    print("== %s  total samples %d" % (name, tot))
    agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
    print("   by reason: " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    body.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
    for r in body[:top]:
        n = int(r[ix["# Samples"]] or 0)
        why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print("   %5.1f%%  %-70s %s" % (100.0 * n / max(tot, 1), r[ix["Source"]].strip()[:70], " ".join("%s:%d" % (w, c) for c, w in why if c)))
    break_after = False
