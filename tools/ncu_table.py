#!/usr/bin/env python
"""Per-launch table from an .ncu-rep (--set full): duration, DRAM bytes and GB/s, tensor-pipe active %,
issue-active %, registers.  `python tools/ncu_table.py rep [hbm_peak_gbs]`"""
import csv, subprocess, sys, io, re, json, os
rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(p))["hbm_gbs"] if os.path.isfile(p) else 6650.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(v, k, d=0.0):
    try:
        return float(v[ix[k]].replace(",", ""))
    except Exception:
        return d
def unit_scale(k, base):
    u = units[ix[k]] if k in ix else ""
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, base)
print("%-58s %9s %8s %9s %9s %7s %7s %7s %5s" % ("kernel", "grid", "us", "dramMB", "GB/s", "%hbm", "tens%", "issue%", "regs"))
for v in vals:
    name = re.sub(r"\(.*", "", v[ix["Kernel Name"]]).replace("void ", "").replace("navc::", "")[:58]
    us = g(v, "gpu__time_duration.sum") * unit_scale("gpu__time_duration.sum", 1.0)
    rd = g(v, "dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum", 1.0)
    wr = g(v, "dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum", 1.0)
    gbs = (rd + wr) / max(us, 1e-9) / 1e3
    print("%-58s %9s %8.1f %9.1f %9.1f %7.1f %7.1f %7.1f %5d" % (
        name, v[ix["launch__grid_size"]], us, (rd + wr) / 1e6, gbs, 100 * gbs / peak,
        g(v, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        g(v, "smsp__issue_active.avg.pct_of_peak_sustained_active"), int(g(v, "launch__registers_per_thread"))))
