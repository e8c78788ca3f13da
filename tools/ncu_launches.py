#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the captured launches)."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1e3
tot = sum(v[1] for v in agg.values())
print("launches %d (skipped first %d), total %.1f us (ncu: cold-cache, serialised -- compare shares)" % (len(rows), skip, tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%10.1f us %5d x %8.1f us %5.1f%%  %s" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k))
