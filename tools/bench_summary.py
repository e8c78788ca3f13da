#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line: value, e2e, per-class roofline table."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f %s  %.2f ms/step   e2e %.0f   precision %s  clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["details"]["precision"], d["clocks"].get("sm_mhz")))
r = d.get("roofline") or {}
for t, c in (r.get("classes") or {}).items():
    print("  %-6s %7.1f us x%3d  share %.3f  useful %.3f issued %s  hbm %.3f  %s" % (
        t, c["us"], c["launches_per_step"], c["share_of_step"], c["frac_useful"], c["frac_issued"], c["frac_hbm"], c["shape"]))
if d.get("parity"):
    for m, p in d["parity"]["modes"].items():
        print("  parity[%s]: %s" % (m, {k: v for k, v in p.items() if k != "differing_videos"}))
