#!/usr/bin/env python
"""Selected raw metrics of every launch in an .ncu-rep:  python tools/ncu_metrics.py rep [regex of metric names]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|sm__pipe_tensor_cycles_active.avg.pct|lts__t_bytes.sum$|lts__throughput.avg.pct|l1tex__throughput.avg.pct|"
                 r"dram__bytes_(read|write).sum$|dram__throughput.avg.pct|smsp__issue_active.avg.pct|sm__throughput.avg.pct|"
                 r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__cycles_active.avg$|sm__cycles_elapsed.max$|"
                 r"lts__t_sectors_srcunit_tex_op_read.sum$|lts__t_sector_hit_rate.pct|smsp__warp_issue_stalled.*_per_warp_active.pct")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
cols = [i for i, h in enumerate(hdr) if pat.search(h)]
for v in vals:
    name = re.sub(r"\(.*", "", v[hdr.index("Kernel Name")]).replace("void ", "").replace("navc::", "")[:60]
    print("== %s grid %s" % (name, v[hdr.index("launch__grid_size")]))
    for i in cols:
        print("   %-85s %14s %s" % (hdr[i], v[i], units[i]))
